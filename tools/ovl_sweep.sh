for cfg in "3 0 0" "2 1 1" "2 2 1" "3 1 1" "1 2 1"; do
  set -- $cfg
  SC_CHUNK_CTAS=$1 SC_LU_CTAS=$2 SC_CHUNK_OVERLAP=$3 python tools/overlap_probe.py 148000 10
done

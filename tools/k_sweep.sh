for k in 5 8 10; do SC_CHUNK_K=$k python tools/pipeline_probe.py 148000 10; done
SC_CHUNK_K=10 SC_CHUNK_SCRATCH_MB=8000 python tools/pipeline_probe.py 148000 10
SC_CHUNK_K=10 python tools/pipeline_probe.py 148000 20
SC_CHUNK_K=20 SC_CHUNK_SCRATCH_MB=8000 python tools/pipeline_probe.py 148000 20

// sc_stage.cuh -- generic-potential stage interface (any Python object implementing the potential protocol).
extern "C" int sc_engine_stage_positions(sc_engine *, int, double, double *, void *) { return fail(SC_ERR_UNSUPPORTED, "stage interface not built yet"); }
extern "C" int sc_engine_stage_apply(sc_engine *, int, double, const double *, const double *, const double *, const double *, double *, void *) { return fail(SC_ERR_UNSUPPORTED, "stage interface not built yet"); }
extern "C" int sc_engine_stage_finish(sc_engine *, double, void *) { return fail(SC_ERR_UNSUPPORTED, "stage interface not built yet"); }

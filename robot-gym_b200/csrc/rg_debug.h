/* Private entry points of librg_cuda.so: development hooks, NOT part of the drop-in boundary (include/rg_cuda.h). */
#ifndef RG_DEBUG_H_
#define RG_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
/* The kernel's own Riccati factor / solve routines on a caller-supplied system (device pointers, float64):
 * k1[6], k2ang[9] (row-major 3x3), k2lin[3]: blocks of the stage cost Q = blkdiag(K2, K1); d[h*21]: packed lower
 * triangles of the D_t; b[6h] -> v_out[6h] with (K^-1 + blkdiag D_t) v = b; flag_out: 1 if a pivot block was not
 * positive definite.  tests/test_gpu_mpc.py checks it against a dense numpy solve. */
int rg_debug_riccati_solve(int horizon, const double* k1, const double* k2ang, const double* k2lin, const double* d,
                           const double* b, double* v_out, int* flag_out, void* stream);
#ifdef RG_DEBUG_TRACE
/* -DRG_DEBUG_TRACE builds only (tools/trace_mpc.py, tools/timeline_mpc.py): per-phase cycle counters of one env. */
int rg_debug_set_trace(double* dev_buf, int env);
#endif
#ifdef __cplusplus
}
#endif
#endif

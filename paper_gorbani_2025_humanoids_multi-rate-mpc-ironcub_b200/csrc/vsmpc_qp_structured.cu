// TEMPORARY forwarding stub — replaced by the structured (sparsity-exploiting) Riccati kernel.
#include "vsmpc_common.cuh"
namespace vsmpc
{
size_t generic_scratch_doubles(const DeviceConfig& cfg);
cudaError_t launch_qp_generic(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                              double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                              int* n_factor, int* n_solve, cudaStream_t s);
size_t structured_scratch_doubles(const DeviceConfig& cfg) { return generic_scratch_doubles(cfg); }
cudaError_t launch_qp_structured(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                 double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                                 int* n_factor, int* n_solve, cudaStream_t s)
{
    return launch_qp_generic(d_cfg, h_cfg, B, qd, ws, scratch, z, st, out_rows, status, n_factor, n_solve, s);
}
} // namespace vsmpc

// Instantiations of the one-launch iteration kernel (admm_pass.cuh), kept in a translation unit
// of their own so that they compile in parallel with the rest of the library.
#include "admm_pass.cuh"
#include "common.cuh"

namespace pogs_b200 {

namespace {
template <typename T, int NV, int B>
void launch_one(unsigned grid, size_t smem, cudaStream_t st, const PassArgs<T>& a, const ParityArgs<T>& par0,
                const ParityArgs<T>& par1, Gate gate, const PeerView& pv, bool pdl) {
  auto kernel = k_admm_pass<T, NV, B>;
  static size_t attr_smem_dev[kMaxDevices] = {};   // per instantiation and device
  size_t& attr_smem = attr_smem_dev[current_device_index()];
  if (attr_smem < smem) {
    POGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int nb = 0;
    POGS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kFusedCta, smem));
    if (nb < 1) throw Error("one-launch iteration kernel does not fit an SM");   // the grid barriers need co-residency
    attr_smem = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kFusedCta); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  POGS_CUDA(cudaLaunchKernelEx(&cfg, kernel, a, par0, par1, gate, pv));
}
}  // namespace

template <typename T>
void launch_admm_pass(int nv, int batch, unsigned grid, size_t smem, cudaStream_t st, const PassArgs<T>& a,
                      const ParityArgs<T>& par0, const ParityArgs<T>& par1, Gate gate, const PeerView& pv, bool pdl) {
#define POGS_AP_CASE(NV, B) \
  case NV * 8 + B: launch_one<T, NV, B>(grid, smem, st, a, par0, par1, gate, pv, pdl); break;
#define POGS_AP_ROW(NV) POGS_AP_CASE(NV, 1) POGS_AP_CASE(NV, 2) POGS_AP_CASE(NV, 4)
  switch (nv * 8 + batch) {
    POGS_AP_ROW(1) POGS_AP_ROW(2) POGS_AP_ROW(3) POGS_AP_ROW(5) POGS_AP_ROW(8)
    default: throw Error("one-launch iteration kernel: no instantiation");
  }
#undef POGS_AP_ROW
#undef POGS_AP_CASE
  POGS_CUDA(cudaGetLastError());
  count_launch();
}

template void launch_admm_pass<float>(int, int, unsigned, size_t, cudaStream_t, const PassArgs<float>&,
                                      const ParityArgs<float>&, const ParityArgs<float>&, Gate, const PeerView&, bool);
template void launch_admm_pass<double>(int, int, unsigned, size_t, cudaStream_t, const PassArgs<double>&,
                                       const ParityArgs<double>&, const ParityArgs<double>&, Gate, const PeerView&, bool);

}  // namespace pogs_b200

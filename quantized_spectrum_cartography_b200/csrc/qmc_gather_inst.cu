// One translation unit per (kernel family, padded rank): compiled with -DQMC_FAMILY=<flat|tiled|lanes as 1|2|3>
// -DQMC_RP=<1|2|4|8|16|32> by build.py so the instantiations build in parallel.
#if QMC_FAMILY == 1
#include "qmc_gather_flat.cuh"
namespace qmc { template int launch_flat_rp<QMC_RP>(const GatherParams&, int, bool, bool, cudaStream_t); }
#elif QMC_FAMILY == 2
#include "qmc_gather_tiled.cuh"
namespace qmc { template int launch_tiled_rp<QMC_RP>(const GatherParams&, int, bool, bool, cudaStream_t); }
#elif QMC_FAMILY == 3
#include "qmc_gather_lanes.cuh"
namespace qmc { template int launch_lanes_rp<QMC_RP>(const GatherParams&, int, bool, bool, cudaStream_t); }
#else
#error "QMC_FAMILY must be 1 (flat), 2 (tiled) or 3 (lanes)"
#endif

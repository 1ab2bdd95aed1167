// explicit instantiation of the BF16 / FP kernels (see iwq_launch.cuh)
#define IWQ_INSTANTIATE
#include "iwq_launch.cuh"
namespace iwq {
IWQ_INSTANTIATE_KIND(IWQ_BF16, KIND_FP)
template cudaError_t launch_col_double<IWQ_BF16>(const LaunchParams&, cudaStream_t);
template cudaError_t launch_seg_double<IWQ_BF16>(const LaunchParams&, cudaStream_t);
}

// explicit instantiation of the F16 / FP kernels (see iwq_launch.cuh)
#define IWQ_INSTANTIATE
#include "iwq_launch.cuh"
namespace iwq {
IWQ_INSTANTIATE_KIND(IWQ_F16, KIND_FP)
template cudaError_t launch_col_double<IWQ_F16>(const LaunchParams&, cudaStream_t);
template cudaError_t launch_seg_double<IWQ_F16>(const LaunchParams&, cudaStream_t);
}

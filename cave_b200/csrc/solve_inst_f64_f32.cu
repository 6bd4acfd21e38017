// Instantiation of the solve kernel: factor precision double, I/O dtype float.
#include "solve_kernel_impl.cuh"
namespace cave {
template cudaError_t launch_solve_t<double, float>(const SolveParams&, int, int, cudaStream_t);
}

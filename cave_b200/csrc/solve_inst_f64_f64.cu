// Instantiation of the solve kernel: factor precision double, I/O dtype double.
#include "solve_kernel_impl.cuh"
namespace cave {
template cudaError_t launch_solve_t<double, double>(const SolveParams&, int, int, cudaStream_t);
}

// Register / spill check of a few K1 instantiations without building the whole library:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -c scripts/dev/ptxas_k1.cu -o /dev/null
#include "../../zvdb_b200/csrc/search_kernel.cuh"
using namespace zvdb;
template __global__ void zvdb::search_layer0_kernel<1, 0, 0, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<1, 0, 1, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<1, 0, 2, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<2, 0, 0, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<2, 0, 1, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<2, 0, 2, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<6, 1, 2, false>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<1, 0, 0, true>(const __grid_constant__ SearchParams);
template __global__ void zvdb::search_layer0_kernel<1, 0, 1, true>(const __grid_constant__ SearchParams);

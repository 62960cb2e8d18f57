// selftest.cu -- exhaustive device-side proofs of the arithmetic shortcuts the kernels take.
//
// hvx_selftest_edge_parameter: the decoupled regular kernel computes the edge parameter d0 / (d0 - d1) with the
// branch-free sequence edge_parameter_int16 (hvx_device.cuh) instead of __fdiv_rn.  Its operands are exact
// integers in [-32768, 32767] (CellWord densities), so the claim "bit-identical" is finite: this kernel checks ALL
// 2^32 operand pairs against edge_parameter (the reference's formula, PV/tests/gpu_transvoxel_emission.rs:305-312,
// with IEEE division) and counts the pairs whose bits differ.
#include <cstring>

#include "hvx_device.cuh"
#include "../../include/hvx.h"

namespace hvx {
namespace {

__global__ void __launch_bounds__(256) edge_parameter_check_kernel(unsigned long long* mismatches, unsigned int* first_bad) {
    // block b covers d0 = b - 32768; threads stride over d1
    const float d0 = static_cast<float>(static_cast<int>(blockIdx.x) - 32768);
    unsigned int bad = 0, witness = 0;
    for (int i = threadIdx.x; i < 65536; i += 256) {
        const float d1 = static_cast<float>(i - 32768);
        const unsigned int want = __float_as_uint(edge_parameter(d0, d1)), got = __float_as_uint(edge_parameter_int16(d0, d1));
        if (want != got) {
            ++bad;
            witness = (blockIdx.x << 16) | static_cast<unsigned int>(i);
        }
    }
    if (bad) {
        atomicAdd(mismatches, static_cast<unsigned long long>(bad));
        atomicExch(first_bad, witness);
    }
}

// every float s in [lo, hi] (consecutive bit patterns): 1 / sqrt(s) by the branch-free sequence vs IEEE sqrt then divide
__global__ void __launch_bounds__(256) inv_sqrt_check_kernel(unsigned int first_bits, unsigned int count, unsigned long long* mismatches,
                                                             unsigned int* first_bad) {
    unsigned int bad = 0, witness = 0;
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < count; i += 256ull * gridDim.x) {
        const unsigned int bits = first_bits + static_cast<unsigned int>(i);
        const float s = __uint_as_float(bits);
        if (__float_as_uint(fdiv(1.0f, fsqrt(s))) != __float_as_uint(inv_sqrt_rn_normal(s))) {
            ++bad;
            witness = bits;
        }
    }
    if (bad) {
        atomicAdd(mismatches, static_cast<unsigned long long>(bad));
        atomicExch(first_bad, witness);
    }
}

}  // namespace
}  // namespace hvx

extern "C" int hvx_selftest_inv_sqrt(int device, uint64_t* mismatches_out, uint32_t* witness_out) {
    if (!mismatches_out) return HVX_E_INVALID_ARGUMENT;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return HVX_E_CUDA;
    unsigned long long* d_count = nullptr;
    unsigned int* d_witness = nullptr;
    int rc = HVX_E_CUDA;
    // the kernel's operand is a sum of three squares, guarded by s > 1e-12: every float from 1e-12 to 2^40
    const float lo = 1.0e-12f, hi = 1099511627776.0f;
    unsigned int lo_bits, hi_bits;
    memcpy(&lo_bits, &lo, 4);
    memcpy(&hi_bits, &hi, 4);
    if (cudaMalloc(&d_count, sizeof *d_count) == cudaSuccess && cudaMalloc(&d_witness, sizeof *d_witness) == cudaSuccess &&
        cudaMemset(d_count, 0, sizeof *d_count) == cudaSuccess && cudaMemset(d_witness, 0, sizeof *d_witness) == cudaSuccess) {
        hvx::inv_sqrt_check_kernel<<<148 * 16, 256>>>(lo_bits, hi_bits - lo_bits + 1u, d_count, d_witness);
        unsigned long long count = 0;
        unsigned int witness = 0;
        if (cudaGetLastError() == cudaSuccess && cudaMemcpy(&count, d_count, sizeof count, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(&witness, d_witness, sizeof witness, cudaMemcpyDeviceToHost) == cudaSuccess) {
            *mismatches_out = count;
            if (witness_out) *witness_out = witness;
            rc = HVX_OK;
        }
    }
    cudaFree(d_count);
    cudaFree(d_witness);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

extern "C" int hvx_selftest_edge_parameter(int device, uint64_t* mismatches_out, uint32_t* witness_out) {
    if (!mismatches_out) return HVX_E_INVALID_ARGUMENT;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return HVX_E_CUDA;
    unsigned long long* d_count = nullptr;
    unsigned int* d_witness = nullptr;
    int rc = HVX_E_CUDA;
    if (cudaMalloc(&d_count, sizeof *d_count) == cudaSuccess && cudaMalloc(&d_witness, sizeof *d_witness) == cudaSuccess &&
        cudaMemset(d_count, 0, sizeof *d_count) == cudaSuccess && cudaMemset(d_witness, 0, sizeof *d_witness) == cudaSuccess) {
        hvx::edge_parameter_check_kernel<<<65536, 256>>>(d_count, d_witness);
        unsigned long long count = 0;
        unsigned int witness = 0;
        if (cudaGetLastError() == cudaSuccess && cudaMemcpy(&count, d_count, sizeof count, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(&witness, d_witness, sizeof witness, cudaMemcpyDeviceToHost) == cudaSuccess) {
            *mismatches_out = count;
            if (witness_out) *witness_out = witness;
            rc = HVX_OK;
        }
    }
    cudaFree(d_count);
    cudaFree(d_witness);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

// The reference's sort call, isolated for the cfg5 microbenchmark: the same
// cub::DeviceRadixSort::SortPairs(begin_bit = 0, end_bit = 32 + bit) on (u64 key, u32 value) pairs that
// /root/reference/src/rasterizer/cuda_rasterizer/rasterizer_impl.cu:419-424 issues, from the CUB that ships with the
// CUDA toolkit the reference is built with here.  TEST / BENCH INFRASTRUCTURE ONLY (part of oracle/_ref).
#include <cub/cub.cuh>
#include <cstdint>

extern "C" int ref_cub_sort_pairs(void* temp, size_t* temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                                  const uint32_t* vals_in, uint32_t* vals_out, int64_t n, int end_bit, void* stream) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit,
                                                (cudaStream_t)stream);
}

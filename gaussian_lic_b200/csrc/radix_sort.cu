// radix_sort.cu -- CUB-free "onesweep" least-significant-digit radix sort of (u64 key, u32 value)
// pairs on key bits [0, end_bit).  Stable.
//
// Replaces cub::DeviceRadixSort::SortPairs(begin_bit = 0, end_bit = 32 + bit) as called at
// reference rasterizer_impl.cu:417-424 (and the Morton sort of simple_knn.cu:210-213).
//
// Structure (8-bit digits, ceil(end_bit/8) passes):
//   1. one histogram kernel builds the global digit histogram of EVERY pass in a single read of
//      the keys (8 B/pair);
//   2. a tiny kernel turns each histogram into exclusive digit offsets;
//   3. one kernel per pass: each CTA ranks a 4096-pair tile (warp-synchronous ranking, stable: a cascade of `bits`
//      ballots -- MATCH.ANY measured ~2x slower on passes whose 256 digits are all populated), resolves its global digit offsets with a per-digit decoupled look-back chain over
//      dynamically ordered CTAs (no second read of the data), stages the tile in shared memory in
//      sorted order and writes digit runs out coalesced (24 B/pair/pass).
// Algorithmic HBM traffic: (8 + 24*passes) B per pair (152 B @ 45 bits).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace glic {

int sort_prepare(int64_t n, int end_bit, void* temp, cudaStream_t s, int digit_bits);
size_t sort_temp_bytes(int64_t n);

namespace {

constexpr int RADIX_BITS = 8;                          // digit width of the 32-bit-key sorts (and of producer-built histograms)
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int RADIX_BITS_MAX = 9;                      // the 64-bit-key sort takes 9-bit digits when that saves a pass (45 bits: 5 instead of 6)
constexpr int RADIX_MAX = 1 << RADIX_BITS_MAX;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS_BIG = 16;                     // 4096 pairs per CTA: the R-sized tile sort
constexpr int SORT_ITEMS_SMALL = 4;                    // 1024 pairs per CTA: inputs that would not fill the GPU with 4096-pair tiles
constexpr int64_t SORT_SMALL_LIMIT = (int64_t)148 * 4 * SORT_THREADS * SORT_ITEMS_BIG;     // below ~2.4 M pairs use the small tile
constexpr int MAX_PASSES = 8;

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_PREFIX = 2u << 30;
constexpr uint32_t VALUE_MASK = (1u << 30) - 1;

struct SortTemp {
    uint32_t* hist;     // [MAX_PASSES][RADIX]  digit histograms -> exclusive offsets
    uint32_t* tickets;  // [MAX_PASSES]
    uint32_t* status;   // [passes][blocks][RADIX]
};

__host__ inline int sort_items(int64_t n) { return n < SORT_SMALL_LIMIT ? SORT_ITEMS_SMALL : SORT_ITEMS_BIG; }
__host__ inline int64_t sort_blocks(int64_t n) { const int64_t tile = (int64_t)SORT_THREADS * sort_items(n); return (n + tile - 1) / tile; }
__host__ inline int64_t sort_blocks_big(int64_t n) { const int64_t tile = (int64_t)SORT_THREADS * SORT_ITEMS_BIG; return (n + tile - 1) / tile; }

__host__ inline SortTemp carve_sort_temp(void* temp) {
    char* p = static_cast<char*>(temp);
    SortTemp t;
    t.hist = reinterpret_cast<uint32_t*>(p);
    t.tickets = t.hist + MAX_PASSES * RADIX_MAX;
    t.status = t.tickets + 32;
    return t;
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- 1. all-pass histogram --------------------------------------------------------------
template <typename KeyT, int RB>
__global__ void __launch_bounds__(256)
sort_histogram_kernel(const KeyT* __restrict__ keys, int64_t n, const unsigned int* __restrict__ n_dev, int passes, int end_bit,
                      uint32_t* __restrict__ hist) {
    constexpr int RADIX_BITS = RB, RADIX = 1 << RB;
    if (n_dev) n = min(n, (int64_t)*n_dev);
    __shared__ uint32_t sh[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const KeyT k = keys[i];
        for (int p = 0; p < passes; ++p) {
            const int shift = p * RADIX_BITS;
            const int bits = min(RADIX_BITS, end_bit - shift);
            const uint32_t d = (uint32_t)(k >> shift) & ((1u << bits) - 1u);
            atomicAdd(&sh[p * RADIX + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x) {
        const uint32_t v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- 2. exclusive scan of each pass's 256 bins --------------------------------------------
__global__ void __launch_bounds__(RADIX_MAX) sort_scan_hist_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_tot[RADIX_MAX / 32];
    uint32_t* h = hist + blockIdx.x * blockDim.x;          // blockDim.x = bins of this sort's digit
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t v = h[tid];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    h[tid] = base + incl - v;
}

// ---- 3. one onesweep pass -------------------------------------------------------------------
template <typename KeyT, int SORT_ITEMS, int RB>
struct __align__(16) PassSmem {
    static constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    static constexpr int RADIX = 1 << RB;
    KeyT keys[SORT_TILE];
    uint32_t vals[SORT_TILE];
    uint32_t warp_cnt[SORT_WARPS][RADIX];
    uint32_t digit_excl[RADIX];
    uint32_t global_base[RADIX];
    uint32_t warp_tot[SORT_WARPS];
    uint32_t block_id;
};

template <typename KeyT, int SORT_ITEMS, int RB>
__global__ void __launch_bounds__(SORT_THREADS, SORT_ITEMS <= 4 ? 6 : (sizeof(KeyT) == 4 ? 5 : 3))
onesweep_pass_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n, const unsigned int* __restrict__ n_dev,
                     int shift, int bits, const uint32_t* __restrict__ digit_base, uint32_t* __restrict__ status,
                     uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
    constexpr int RADIX_BITS = RB, RADIX = 1 << RB;
    constexpr int DPT = RADIX / SORT_THREADS;               // digits owned by one thread in the per-digit phase (1 or 2)
    PassSmem<KeyT, SORT_ITEMS, RB>& sm = *reinterpret_cast<PassSmem<KeyT, SORT_ITEMS, RB>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dmask = (1u << bits) - 1u;

    if (tid == 0) sm.block_id = atomicAdd(ticket, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&sm.warp_cnt[0][0])[i] = 0;
    __syncthreads();
    if (n_dev) n = min(n, (int64_t)*n_dev);   // capacity-sized launch: the true count lives on the device
    const int64_t block = sm.block_id;
    const int64_t tile_base = block * SORT_TILE;
    if (tile_base >= n) return;                 // (CTA-uniform) nothing for this ticket
    const int count = (int)min((int64_t)SORT_TILE, n - tile_base);

    // -- load (warp-striped: lane l, item j <-> element warp*512 + j*32 + l; order-preserving)
    KeyT k[SORT_ITEMS];
    uint32_t rank[SORT_ITEMS];
    const int wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const int li = wbase + j * 32 + lane;
        k[j] = li < count ? keys_in[tile_base + li] : (KeyT)~(KeyT)0;
    }
    // -- stable rank inside the warp's digit streams
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const int li = wbase + j * 32 + lane;
        const bool valid = li < count;
        const uint32_t d = (uint32_t)(k[j] >> shift) & dmask;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        uint32_t r = 0;
        uint32_t bpeers = vmask;
        {   // peers from `bits` ballots: cost independent of the digit distribution (MATCH.ANY stalls ~2x longer on
            // passes whose 256 digits are all populated; ncu: profiles/r01_prof_sort_raw.csv)
#pragma unroll
            for (int b = 0; b < RADIX_BITS; ++b) {
                if (b < bits) {
                    const bool bit = (d >> b) & 1u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit && valid);
                    bpeers &= bit ? bal : ~bal;
                }
            }
        }
        if (valid) {
            const uint32_t peers = bpeers;
            const int leader = __ffs(peers) - 1;
            uint32_t c = 0;
            if (lane == leader) {
                c = sm.warp_cnt[warp][d];
                sm.warp_cnt[warp][d] = c + __popc(peers);
            }
            c = __shfl_sync(peers, c, leader);
            r = c + __popc(peers & lt_mask);
        }
        rank[j] = r;
        __syncwarp();
    }
    __syncthreads();

    // -- per digit (thread t owns digits t*DPT .. t*DPT + DPT-1): warp bases, CTA total, look-back
    {
        uint32_t tot[DPT];
        uint32_t mine = 0;
#pragma unroll
        for (int e = 0; e < DPT; ++e) {
            const int d = tid * DPT + e;
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; ++w) {
                const uint32_t c = sm.warp_cnt[w][d];
                sm.warp_cnt[w][d] = t;
                t += c;
            }
            tot[e] = t;
            mine += t;
        }
        // CTA-local exclusive scan over digits (placement inside the staged tile)
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sm.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t dbase = 0;
        for (int w = 0; w < warp; ++w) dbase += sm.warp_tot[w];
        uint32_t excl = dbase + incl - mine;
#pragma unroll
        for (int e = 0; e < DPT; ++e) {
            const int d = tid * DPT + e;
            sm.digit_excl[d] = excl;
            // decoupled look-back for this digit
            uint32_t* my = status + (size_t)block * RADIX + d;
            uint32_t prev = 0;
            if (block == 0) {
                st_volatile(my, FLAG_PREFIX | tot[e]);
            } else {
                st_volatile(my, FLAG_AGG | tot[e]);
                int64_t b = block - 1;
                unsigned spins = 0;
                while (true) {
                    const uint32_t s = ld_volatile(status + (size_t)b * RADIX + d);
                    const uint32_t f = s & ~VALUE_MASK;
                    if (f == 0) {
                        if (++spins > (1u << 26)) __trap();   // fail loudly rather than hang the GPU
                        continue;
                    }
                    prev += s & VALUE_MASK;
                    if (f == FLAG_PREFIX) break;
                    --b;
                }
                st_volatile(my, FLAG_PREFIX | (prev + tot[e]));
            }
            sm.global_base[d] = digit_base[d] + prev - excl;
            excl += tot[e];
        }
    }
    __syncthreads();

    // -- stage the tile in sorted order
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        const int li = wbase + j * 32 + lane;
        if (li < count) {
            const uint32_t d = (uint32_t)(k[j] >> shift) & dmask;
            const uint32_t pos = sm.digit_excl[d] + sm.warp_cnt[warp][d] + rank[j];
            sm.keys[pos] = k[j];
            sm.vals[pos] = vals_in[tile_base + li];
        }
    }
    __syncthreads();

    // -- coalesced write-out of digit runs
    for (int i = tid; i < count; i += SORT_THREADS) {
        const KeyT key = sm.keys[i];
        const uint32_t d = (uint32_t)(key >> shift) & dmask;
        const size_t dst = (size_t)sm.global_base[d] + i;
        keys_out[dst] = key;
        vals_out[dst] = sm.vals[i];
    }
}

}  // namespace

// zeroes the histograms, tickets and look-back status words of a sort of n pairs on bits [0, end_bit)
int sort_prepare(int64_t n, int end_bit, void* temp, cudaStream_t s, int digit_bits) {
    if (n <= 0) return GLIC_OK;
    const int passes = (end_bit + digit_bits - 1) / digit_bits;
    const size_t used = sizeof(uint32_t) * (MAX_PASSES * RADIX_MAX + 32 + (size_t)passes * sort_blocks(n) * ((size_t)1 << digit_bits));
    GLIC_CUDA_TRY(cudaMemsetAsync(temp, 0, used, s));
    return GLIC_OK;
}

uint32_t* sort_hist(void* temp) { return static_cast<uint32_t*>(temp); }      // [MAX_PASSES][256], pass-major

size_t sort_temp_bytes(int64_t n) {
    const size_t blocks = (size_t)sort_blocks(n > 0 ? n : 1);
    return sizeof(uint32_t) * (MAX_PASSES * RADIX_MAX + 32 + (size_t)MAX_PASSES * blocks * RADIX_MAX) + 256;
}

template <typename KeyT, int ITEMS, int RB>
static int run_passes(int64_t n, int passes, int end_bit, int64_t blocks, KeyT* keys[2], uint32_t* vals[2], const SortTemp& t, cudaStream_t s,
                      const unsigned int* n_dev) {
    constexpr int BINS = 1 << RB;
    static bool attr_set[64] = {};
    if (first_use_on_device(attr_set)) {
        GLIC_CUDA_TRY(cudaFuncSetAttribute(onesweep_pass_kernel<KeyT, ITEMS, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(PassSmem<KeyT, ITEMS, RB>)));
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * RB;
        const int bits = min(RB, end_bit - shift);
        onesweep_pass_kernel<KeyT, ITEMS, RB><<<(unsigned)blocks, SORT_THREADS, sizeof(PassSmem<KeyT, ITEMS, RB>), s>>>(
            keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, n_dev, shift, bits, t.hist + p * BINS,
            t.status + (size_t)p * blocks * BINS, t.tickets + p);
        GLIC_LAUNCH_CHECK();
        cur ^= 1;
    }
    return cur;
}

// Digit width of a sort.  The pass kernel is generic (8- or 9-bit digits; GLIC_SORT_DIGIT_BITS=9 selects the wider one for
// 64-bit keys when it saves a pass), but 8 stays the default: at cfg5 (50 M pairs, 45 bits) 5 passes of 9 bits measured
// 3.96 ms against 3.78 ms for 6 passes of 8 bits -- a pass is bound by the warp-synchronous ranking (one ballot per digit bit
// and item) and by the per-digit look-back, not by its 24 B/pair of HBM traffic, so a wider digit costs more than the saved
// pass returns (profiles/README.md, sort section).
template <typename KeyT>
static inline int sort_digit_bits(int end_bit) {
    static const int want = getenv("GLIC_SORT_DIGIT_BITS") ? atoi(getenv("GLIC_SORT_DIGIT_BITS")) : 8;
    if (want == 9 && sizeof(KeyT) == 8 && (end_bit + 8) / 9 < (end_bit + 7) / 8) return 9;
    return 8;
}

// hist_ready: the caller zeroed the temp block with sort_prepare() and a producer kernel already accumulated the digit
// histograms of every pass into sort_hist(temp) (8-bit digits): the histogram kernel -- one more read of all keys -- is skipped.
template <typename KeyT>
static int launch_sort_pairs_t(int64_t n, int end_bit, KeyT* keys[2], uint32_t* vals[2], void* temp, size_t temp_bytes,
                               cudaStream_t s, const unsigned int* n_dev = nullptr, bool hist_ready = false) {
    if (end_bit < 1 || end_bit > (int)(8 * sizeof(KeyT))) { set_error("sort: end_bit out of range"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (n >= (int64_t)VALUE_MASK) { set_error("sort: n too large"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (n <= 0) return 0;
    const int rb = hist_ready ? 8 : sort_digit_bits<KeyT>(end_bit);
    const int passes = (end_bit + rb - 1) / rb;
    if (passes > MAX_PASSES) { set_error("sort: too many passes"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (temp_bytes < sort_temp_bytes(n)) { set_error("sort: temp too small"); return GLIC_ERR_WORKSPACE; }
    const int64_t blocks = sort_blocks(n);
    SortTemp t = carve_sort_temp(temp);
    if (!hist_ready) {
        if (int e = sort_prepare(n, end_bit, temp, s, rb)) return e;
        int hist_blocks = (int)std::min<int64_t>((n + 256 * 16 - 1) / (256 * 16), (int64_t)148 * 8);
        if (rb == 9) sort_histogram_kernel<KeyT, 9><<<hist_blocks, 256, 0, s>>>(keys[0], n, n_dev, passes, end_bit, t.hist);
        else sort_histogram_kernel<KeyT, 8><<<hist_blocks, 256, 0, s>>>(keys[0], n, n_dev, passes, end_bit, t.hist);
        GLIC_LAUNCH_CHECK();
    }
    sort_scan_hist_kernel<<<passes, 1 << rb, 0, s>>>(t.hist);
    GLIC_LAUNCH_CHECK();
    if (rb == 9) {
        if constexpr (sizeof(KeyT) == 8) return run_passes<KeyT, SORT_ITEMS_BIG, 9>(n, passes, end_bit, sort_blocks_big(n), keys, vals, t, s, n_dev);
    }
    return sort_items(n) == SORT_ITEMS_SMALL ? run_passes<KeyT, SORT_ITEMS_SMALL, 8>(n, passes, end_bit, blocks, keys, vals, t, s, n_dev)
                                             : run_passes<KeyT, SORT_ITEMS_BIG, 8>(n, passes, end_bit, blocks, keys, vals, t, s, n_dev);
}

int launch_sort_pairs(int64_t n, int end_bit, uint64_t* keys[2], uint32_t* vals[2], void* temp, size_t temp_bytes,
                      cudaStream_t s) {
    return launch_sort_pairs_t<uint64_t>(n, end_bit, keys, vals, temp, temp_bytes, s);
}

int launch_sort_pairs32(int64_t n, int end_bit, uint32_t* keys[2], uint32_t* vals[2], void* temp, size_t temp_bytes,
                        cudaStream_t s, const unsigned int* n_dev, bool hist_ready) {
    return launch_sort_pairs_t<uint32_t>(n, end_bit, keys, vals, temp, temp_bytes, s, n_dev, hist_ready);
}

}  // namespace glic

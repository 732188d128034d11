// NatureCNN on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), bf16 operands,
// fp32 accumulation.  Reference network: cleanrl/ppo_atari_envpool.py:123-139.
//
// Data layout in HBM (all activations are LINEAR PIXEL GRIDS: row = grid position, 128 bytes = 64 channels)
//   frames   bf16 [n, 21x21, 64]   space-to-depth(4) of the uint8 frames: channel = c*16 + sy*4 + sx of source
//                                  pixel (4Y+sy, 4X+sx); written once per env step (tc_frames_to_s2d), 0..255 is
//                                  exact in bf16 and the /255 of ppo_atari_envpool.py:144 is applied to the fp32
//                                  accumulator.  conv1 (8x8 s4) = 2x2 stride-1 conv on this grid.
//   act1     bf16 [n, 10x10, 128]  conv1 output as 2x2 cells (space-to-depth(2)): conv2 (4x4 s2) = 2x2 stride-1.
//   act2     bf16 [n,  9x9,  64]   act3 bf16 [n, 7x7, 64]   hidden bf16 [n, 512]
//   gradients of act3 / act2 are written twice by their producer: on the consumer-weight-gradient's grid
//   (zeros at positions that are not valid outputs) and zero-padded for the data-gradient's "full" correlation;
//   d(act1) lives on the 21x21 grid with 32 channels.  Never-written positions rely on a zero-initialised
//   workspace.  The minibatch gather b_obs[mb_inds] (ppo.py:250, 3.7 GB fp32 per minibatch in the reference)
//   is an image-index indirection inside the conv1 kernels.
//
// Kernels
//   tc_conv_win<BN,CPR,STAGES,NTAPS>   stride-1 "window" convolution (conv1/2/3 forward, conv3/conv2 data-gradient):
//       GEMM rows enumerate grid positions, so tap (dy,dx) of row r is row r + dy*Wp + dx.  A persistent CTA
//       stages ONE window of 128+maxshift rows per tile as a TMA box (cp.async.bulk.tensor; for conv1 a 3-D box
//       whose image coordinate is the minibatch gather) and every tap is a UMMA descriptor whose start address is
//       shifted by whole 128-byte rows (legal for SWIZZLE_128B: the pattern is a function of the smem address
//       bits; tools/experiments/umma_shift_test.cu).  Weights stay resident in smem.  Warp 0 = TMA producer, warp 1 issues tcgen05.mma
//       (warp-uniform loop, elected lane); warps 2-9 = two epilogue groups that own alternate tiles and drain
//       double-buffered accumulators; ReLU masks are exchanged between forward and backward as bits.
//   tc_wgrad_win                 conv weight gradients: dW^T[(tap,c), co] = sum_r X[r+shift_tap, c] * dY[r, co]; the same
//       row images are read as MN-major operands (rows = reduction index), taps again by row shifts; the bias
//       gradient (column sums of dY) is accumulated by the four dY warps from the staged tiles.
//   tc_gemm_tma<BN,STAGES>       fc forward / data-gradient: both operands are TMA boxes of row-major matrices.
//   tc_wgrad_tma                 fc weight gradient (MN-major views of TMA-loaded dhid / act3 row boxes).
//   tc_heads_*                   the A+1 head outputs in fp32 on CUDA cores.
#include <cuda.h>            // CUtensorMap types only; the encoder is resolved at run time (no libcuda link)
#include "common.cuh"
#include "tc_common.cuh"

namespace b200rl {
using namespace tc;
typedef __nv_bfloat16 bf16;

// Plain GEMM out[M, N] = A[M, 64*nchunks] . Bw[N, 64*nchunks]^T with a fused epilogue (tc_gemm_tma: fc layer)
struct KGemmParams {
    const void* A;         // row-major bf16 [M, 64*nchunks]
    int64_t M;
    int nchunks;           // K = 64*nchunks
    const bf16* Bw;        // packed weights [N, 64*nchunks]
    int N;
    // ---- epilogue
    bf16* out;
    int ldo;
    const float* bias;
    float scale;
    int relu;
    // ReLU masks as bits, word (row * N/32 + col/32) of a dense [M, N] tensor (see WinParams)
    const uint32_t* mask_bits;   // multiply the output by the mask (data-gradient)
    uint32_t* mask_out;          // record (output > 0) (forward with relu)
    // fc data-gradient only: write dact3 on the 9x9 linear grid (out) and zero-padded 11x11 grid (out2)
    int dual_dact3;
    bf16* out2;
};

__device__ __forceinline__ int4 ldg16(const void* p) { return __ldg(reinterpret_cast<const int4*>(p)); }

// ------------------------------------------------------------------ kernel 1c: "window" convolution
// Stride-1 convolutions over activations stored as a LINEAR pixel grid [n*G rows, CPR*64 channels]
// (G = Hp*Wp grid positions per image).  GEMM rows enumerate grid positions, so tap (dy,dx) of a row is
// simply the row `dy*Wp+dx` further down: the CTA stages ONE window of 128+maxshift rows per tile and
// every tap is a tcgen05 operand descriptor whose start address is shifted by whole 128-byte rows
// (the SWIZZLE_128B pattern is a function of the shared-memory address bits, so any row shift is legal:
// tools/experiments/umma_shift_test.cu).  Each activation row is therefore read from L2 once per tile
// instead of once per tap, and the producers do no im2col index arithmetic at all.  Grid positions whose
// window would leave the image (X >= vW or Y >= vH) are computed but not stored.
enum { WOUT_DENSE = 0, WOUT_S2D2 = 1, WOUT_DACT2 = 2, WOUT_DACT1 = 3 };
struct WinParams {
    const bf16* A;           // [n*G, CPR*64]
    const int64_t* rows;     // optional image gather (conv1 reads the rollout through mb_inds)
    int64_t M;               // n*G
    int n, G, Wp;
    // image-aligned tiling (conv1): every image owns 2^tpi_shift tiles of 128 grid rows (rows >= G are padding),
    // so a window never spans two images and the minibatch gather is just the TMA box's image coordinate.
    // 0 = tiles walk the linear grid [n*G] (activations produced by this library, always contiguous).
    int tpi_shift;
    int64_t n_images;        // images addressable through `rows` (size of the tensor map's outer dimension)
    int ntaps;
    int shift[16];           // dy*Wp + dx per tap (non-negative)
    int WR;                  // window rows: 128 + max shift, rounded up to 8
    const bf16* Bw;          // packed weights [N][ntaps*CPR*64]
    int N;
    int vH, vW;              // valid outputs: Y < vH && X < vW
    int out_mode;
    bf16* out;               // primary output
    bf16* out2;              // WOUT_DACT2: padded 11x11 copy
    // ReLU masks travel as BITS (1 = the forward activation was > 0), one 32-bit word per 32 channels, in the row
    // order of the tensor they describe: 16x fewer bytes than re-reading the bf16 activation
    const uint32_t* mask_bits;   // input mask (data-gradient kernels): words of the row this thread writes
    uint32_t* mask_out;          // output mask (forward kernels with relu)
    const float* bias;
    float scale;
    int relu;
};

template <int BN, int CPR, int STAGES, int NTAPS>
__global__ void __launch_bounds__(320, 1) tc_conv_win(const __grid_constant__ CUtensorMap tmA, const WinParams p,
                                                      int total_tiles) {
    constexpr int B_CHUNK = BN * 128;
    constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nchunks = p.ntaps * CPR;
    const int K = nchunks * 64;
    const int IMG = p.WR * 128;                 // one 64-channel column image of the window
    const int STAGE_BYTES = IMG * CPR;
    uint8_t* sW = smem;
    uint8_t* sRing = smem + (size_t)nchunks * B_CHUNK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    for (int idx = tid; idx < nchunks * BN * 8; idx += blockDim.x) {
        const int c16 = idx & 7;
        int t = idx >> 3;
        const int r = t % BN; const int j = t / BN;
        int4 v = make_int4(0, 0, 0, 0);
        if (r < p.N) v = ldg16(p.Bw + (int64_t)r * K + j * 64 + c16 * 8);
        *reinterpret_cast<int4*>(sW + (size_t)j * B_CHUNK + img_off(r, c16)) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_smem;
    // each CTA walks a CONTIGUOUS range of tiles: with the minibatch gather every image (3-4 tiles) is then
    // touched by one SM only (TLB / L2 locality), and the image indices of tile+1 can be prefetched
    const int tile_begin = (int)(((int64_t)total_tiles * blockIdx.x) / gridDim.x);
    const int tile_end = (int)(((int64_t)total_tiles * (blockIdx.x + 1)) / gridDim.x);

    if (warp == 0) {
        // ======================= TMA producer: the window is one rectangular box per 64-channel column chunk ====
        if (tid == 0) {
            uint32_t q = 0;
            const int tmask = (1 << p.tpi_shift) - 1;
            // image-aligned mode: the box's image coordinate is the (optional) minibatch gather; the index of the
            // NEXT tile's image is fetched one tile ahead so the dependent load never delays a TMA issue
            int z_next = 0;
            if (p.tpi_shift && tile_begin < tile_end) {
                const int img = tile_begin >> p.tpi_shift;
                z_next = p.rows ? (int)__ldg(p.rows + img) : img;
            }
            for (int tile = tile_begin; tile < tile_end; ++tile, ++q) {
                const uint32_t s = q % STAGES;
                const int z = z_next;
                if (p.tpi_shift && tile + 1 < tile_end) {
                    const int img = (tile + 1) >> p.tpi_shift;
                    z_next = p.rows ? (int)__ldg(p.rows + img) : img;
                }
                if (q >= (uint32_t)STAGES) mbar_wait(&empty_bar[s], ((q / STAGES) - 1) & 1);
                const uint32_t dst = smem_u32(sRing + (size_t)s * STAGE_BYTES);
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)STAGE_BYTES);
                if (p.tpi_shift) {
#pragma unroll
                    for (int c = 0; c < CPR; ++c) tma_load_3d(dst + c * IMG, &tmA, c * 64, (tile & tmask) * 128, z, &full_bar[s]);
                } else {
#pragma unroll
                    for (int c = 0; c < CPR; ++c) tma_load_2d(dst + c * IMG, &tmA, c * 64, tile * 128, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer: the WHOLE warp walks the tile loop (uniform control flow keeps the
        // descriptor arithmetic in uniform registers), one elected lane issues.  Descriptors differ from a per-stage /
        // per-tap base only in their 14-bit start-address field, so each MMA costs two 32-bit adds.
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        const uint64_t desc_hi = desc_kmajor(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo_flags = (uint32_t)(desc_kmajor(0) & 0xFFFFFFFFull);       // LBO field
        const uint32_t w_lo = ((smem_u32(sW) & 0x3FFFFu) >> 4) | desc_lo_flags;
        uint32_t q = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile, ++q) {
            const uint32_t acc = q & 1, s = q % STAGES;
            if (q >= 2) mbar_wait(&tempty_bar[acc], ((q >> 1) - 1) & 1);
            mbar_wait(&full_bar[s], (q / STAGES) & 1);
            tc_fence_after_sync();
            if (leader) {
                const uint32_t d_addr = tmem_d + acc * BN;
                const uint32_t win_lo = ((smem_u32(sRing + (size_t)s * STAGE_BYTES) & 0x3FFFFu) >> 4) | desc_lo_flags;
#pragma unroll
                for (int t = 0; t < NTAPS; ++t) {
#pragma unroll
                    for (int c = 0; c < CPR; ++c) {
                        const uint32_t a_lo = win_lo + (uint32_t)((c * IMG) >> 4) + (uint32_t)p.shift[t] * 8u;
                        const uint32_t b_lo = w_lo + (uint32_t)(((t * CPR + c) * B_CHUNK) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_bf16(d_addr, desc_hi | (uint64_t)(a_lo + 2 * kk), desc_hi | (uint64_t)(b_lo + 2 * kk), idesc,
                                      (t | c | kk) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue: warps 2-9 = two groups of four (one warp per TMEM lane quadrant).
        // Group h owns accumulator buffer h, i.e. every other tile of the CTA's range, and handles ALL BN columns
        // of its rows: the row -> (image, Y, X) -> output offset arithmetic is paid once per tile and thread, and
        // is incremental (a tile step is 256 grid rows; Y = rem / Wp by multiply-shift, exact for G*Wp < 65536).
        const int ew = warp & 3;
        const int h = (warp - 2) >> 2;
        const int lrow = ew * 32 + (tid & 31);
        const uint32_t mW = (65536u + (uint32_t)p.Wp - 1u) / (uint32_t)p.Wp;
        const int step_i = 256 / p.G, step_rem = 256 % p.G;
        int64_t r = ((int64_t)tile_begin + h) * 128 + lrow;      // linear-grid mode state (unused when image-aligned)
        int i_lin = (int)(r / p.G);
        int rem_lin = (int)(r - (int64_t)i_lin * p.G);
        const int tmask = (1 << p.tpi_shift) - 1;
        const uint32_t lane_addr = tmem_d + h * BN + ((uint32_t)(ew * 32) << 16);
        constexpr int NW = BN / 32;                          // 32-column groups = mask words per row
        uint32_t k = 0;                                      // use count of accumulator buffer h
        for (int tile = tile_begin + h; tile < tile_end; tile += 2, ++k) {
            int i = i_lin, rem = rem_lin;
            bool inside = r < p.M;
            if (p.tpi_shift) {                               // image-aligned tiles: rows >= G of an image are padding
                i = tile >> p.tpi_shift;
                rem = ((tile & tmask) << 7) + lrow;
                inside = rem < p.G;
            }
            const int Y = (int)(((uint32_t)rem * mW) >> 16), X = rem - Y * p.Wp;
            const bool valid = inside && (Y < p.vH) && (X < p.vW);
            int64_t o1 = 0, o2 = 0, ob = 0;
            if (p.out_mode == WOUT_DENSE) {
                const int64_t orow = ((int64_t)i * p.vH + Y) * p.vW + X;
                o1 = orow * p.N; ob = orow * (p.N >> 5);
            } else if (p.out_mode == WOUT_S2D2) {
                const int64_t cell = ((int64_t)i * 10 + (Y >> 1)) * 10 + (X >> 1);
                const int cls = (Y & 1) * 2 + (X & 1);
                o1 = cell * 128 + cls * 32; ob = cell * 4 + cls;
            } else if (p.out_mode == WOUT_DACT2) {
                o1 = ((int64_t)i * 100 + Y * 10 + X) * 64;                 // 10-grid linear (conv2 wgrad)
                o2 = ((int64_t)i * 121 + (Y + 1) * 11 + (X + 1)) * 64;     // zero-padded 11x11 (conv2 dgrad)
                ob = ((int64_t)i * 81 + Y * 9 + X) * 2;                    // act2 mask words
            } else {
                ob = ((int64_t)i * 100 + Y * 10 + X) * 4;                  // act1 (2x2 cells) mask words
            }
            // the row's mask words are requested BEFORE waiting for the accumulator (latency overlaps the MMAs)
            uint32_t mb[NW];
#pragma unroll
            for (int g = 0; g < NW; ++g) mb[g] = 0xFFFFFFFFu;
            if (p.mask_bits != nullptr && valid) {
                if (NW == 4) {
                    const int4 t = ldg16(p.mask_bits + ob);
                    mb[0] = (uint32_t)t.x; mb[1 % NW] = (uint32_t)t.y; mb[2 % NW] = (uint32_t)t.z; mb[3 % NW] = (uint32_t)t.w;
                } else if (NW == 2) {
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p.mask_bits + ob));
                    mb[0] = t.x; mb[1 % NW] = t.y;
                } else {
                    mb[0] = __ldg(p.mask_bits + ob);
                }
            }
            mbar_wait(&tfull_bar[h], k & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int g = 0; g < NW; ++g) {
                uint32_t v[32];
                tmem_ld32(lane_addr + g * 32, v);
                tmem_ld_wait();
                if (g == NW - 1) {             // accumulator drained: hand the buffer back before the global stores
                    tc_fence_before_sync();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&tempty_bar[h]);
                }
                if (!valid || g * 32 >= p.N) continue;
                if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + g * 32);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 bv = __ldg(bp + e);
                        v[4 * e] = __float_as_uint(fmaf(__uint_as_float(v[4 * e]), p.scale, bv.x));
                        v[4 * e + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 1]), p.scale, bv.y));
                        v[4 * e + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 2]), p.scale, bv.z));
                        v[4 * e + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 3]), p.scale, bv.w));
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * p.scale);
                }
                if (p.relu) {
                    uint32_t bits = 0u;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float f = __uint_as_float(v[e]);
                        const bool pos = f > 0.f;
                        bits |= (pos ? 1u : 0u) << e;
                        v[e] = pos ? v[e] : 0u;
                    }
                    if (p.mask_out) p.mask_out[ob + g] = bits;
                }
                if (p.mask_bits) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) if (!((mb[g] >> e) & 1u)) v[e] = 0u;
                }
                int4 w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    w[e].x = (int)pack_bf16x2(__uint_as_float(v[8 * e]), __uint_as_float(v[8 * e + 1]));
                    w[e].y = (int)pack_bf16x2(__uint_as_float(v[8 * e + 2]), __uint_as_float(v[8 * e + 3]));
                    w[e].z = (int)pack_bf16x2(__uint_as_float(v[8 * e + 4]), __uint_as_float(v[8 * e + 5]));
                    w[e].w = (int)pack_bf16x2(__uint_as_float(v[8 * e + 6]), __uint_as_float(v[8 * e + 7]));
                }
                int4* dst;
                if (p.out_mode == WOUT_DACT1) {
                    // column group g = (py,px) of the cell -> input pixel (2Y+py, 2X+px) of the 21-grid, 32 channels
                    dst = reinterpret_cast<int4*>(p.out + ((int64_t)i * 441 + (2 * Y + (g >> 1)) * 21 + 2 * X + (g & 1)) * 32);
                } else {
                    dst = reinterpret_cast<int4*>(p.out + o1 + g * 32);
                }
                dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
                if (p.out_mode == WOUT_DACT2) {
                    int4* dst2 = reinterpret_cast<int4*>(p.out2 + o2 + g * 32);
                    dst2[0] = w[0]; dst2[1] = w[1]; dst2[2] = w[2]; dst2[3] = w[3];
                }
            }
            r += 256; i_lin += step_i; rem_lin += step_rem;
            if (rem_lin >= p.G) { rem_lin -= p.G; ++i_lin; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}


// ------------------------------------------------------------------ kernel 1d: TMA-fed GEMM (fc forward / data-gradient)
// Plain row-major operands => the tiles are rectangular boxes: ONE thread issues cp.async.bulk.tensor (TMA,
// SWIZZLE_128B) loads for the A chunk [128 x 64] and the weight chunk [BN x 64]; the hardware does the address
// generation, zero-fills out-of-range rows and signals the stage's mbarrier with complete_tx.  Warp 0 = TMA
// producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue over double-buffered accumulators.
template <int BN, int STAGES>
__global__ void __launch_bounds__(320, 1) tc_gemm_tma(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB,
                                                      const KGemmParams p, int total_tiles, int ntiles_n) {
    constexpr int A_BYTES = 128 * 128;
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = (2 * BN) < 32 ? 32 : 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nch = p.nchunks;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, TMEM_COLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_smem;

    if (warp == 0) {
        if ((tid & 31) == 0) {
            uint32_t q = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / ntiles_n, n0 = (tile - mt * ntiles_n) * BN;
                for (int j = 0; j < nch; ++j, ++q) {
                    const uint32_t s = q % STAGES;
                    if (q >= (uint32_t)STAGES) mbar_wait(&empty_bar[s], ((q / STAGES) - 1) & 1);
                    const uint32_t dst = smem_u32(smem + (size_t)s * STAGE_BYTES);
                    mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(dst, &tmA, j * 64, mt * 128, &full_bar[s]);
                    tma_load_2d(dst + A_BYTES, &tmB, j * 64, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: whole warp walks the loop (uniform control flow), one elected lane issues
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        const uint64_t desc_hi = desc_kmajor(0) & 0xFFFFFFFF00000000ull;
        const uint32_t lo_flags = (uint32_t)(desc_kmajor(0) & 0xFFFFFFFFull);
        uint32_t q = 0, t = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++t) {
            const uint32_t acc = t & 1;
            if (t >= 2) mbar_wait(&tempty_bar[acc], ((t >> 1) - 1) & 1);
            tc_fence_after_sync();
            const uint32_t d_addr = tmem_d + acc * BN;
            for (int j = 0; j < nch; ++j, ++q) {
                const uint32_t s = q % STAGES;
                mbar_wait(&full_bar[s], (q / STAGES) & 1);
                tc_fence_after_sync();
                if (leader) {
                    const uint32_t stage_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
                    const uint32_t a_lo = ((stage_addr & 0x3FFFFu) >> 4) | lo_flags;
                    const uint32_t b_lo = (((stage_addr + A_BYTES) & 0x3FFFFu) >> 4) | lo_flags;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(d_addr, desc_hi | (uint64_t)(a_lo + 2 * kk), desc_hi | (uint64_t)(b_lo + 2 * kk), idesc,
                                  (j | kk) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                }
                __syncwarp();
            }
            if (leader) umma_commit(&tfull_bar[acc]);
            __syncwarp();
        }
    } else {
        // warps 2-9 = two groups of four (one warp per TMEM lane quadrant); group h owns accumulator buffer h
        // (every other tile of this CTA) and handles all BN columns of its rows in 32-column steps
        const int ew = warp & 3;
        const int h = (warp - 2) >> 2;
        const int lrow = ew * 32 + (tid & 31);
        const uint32_t lane_addr = tmem_d + h * BN + ((uint32_t)(ew * 32) << 16);
        const int nwords = p.N >> 5;                         // mask words per row (N is a multiple of 32)
        constexpr int NW = BN / 32;
        uint32_t k = 0;
        for (int tile = blockIdx.x + h * (int)gridDim.x; tile < total_tiles; tile += 2 * (int)gridDim.x, ++k) {
            const int mt = tile / ntiles_n, n0 = (tile - mt * ntiles_n) * BN;
            const int r = mt * 128 + lrow;
            const bool rvalid = r < (int)p.M;
            const int64_t ooff = (int64_t)r * p.ldo;
            const int64_t wb = (int64_t)r * nwords + (n0 >> 5);
            uint32_t mb[NW];
#pragma unroll
            for (int g = 0; g < NW; ++g)
                mb[g] = (p.mask_bits != nullptr && rvalid && n0 + g * 32 < p.N) ? __ldg(p.mask_bits + wb + g) : 0xFFFFFFFFu;
            mbar_wait(&tfull_bar[h], k & 1);
            tc_fence_after_sync();
#pragma unroll
            for (int g = 0; g < NW; ++g) {
                uint32_t v[32];
                tmem_ld32(lane_addr + g * 32, v);
                tmem_ld_wait();
                if (g == NW - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&tempty_bar[h]);
                }
                const int col = n0 + g * 32;
                if (!rvalid || col >= p.N) continue;
                if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 bv = __ldg(bp + e);
                        v[4 * e] = __float_as_uint(fmaf(__uint_as_float(v[4 * e]), p.scale, bv.x));
                        v[4 * e + 1] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 1]), p.scale, bv.y));
                        v[4 * e + 2] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 2]), p.scale, bv.z));
                        v[4 * e + 3] = __float_as_uint(fmaf(__uint_as_float(v[4 * e + 3]), p.scale, bv.w));
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * p.scale);
                }
                if (p.relu) {
                    uint32_t bits = 0u;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const bool pos = __uint_as_float(v[e]) > 0.f;
                        bits |= (pos ? 1u : 0u) << e;
                        v[e] = pos ? v[e] : 0u;
                    }
                    if (p.mask_out) p.mask_out[wb + g] = bits;
                }
                if (p.mask_bits) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) if (!((mb[g] >> e) & 1u)) v[e] = 0u;
                }
                int4 w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    w[e].x = (int)pack_bf16x2(__uint_as_float(v[8 * e]), __uint_as_float(v[8 * e + 1]));
                    w[e].y = (int)pack_bf16x2(__uint_as_float(v[8 * e + 2]), __uint_as_float(v[8 * e + 3]));
                    w[e].z = (int)pack_bf16x2(__uint_as_float(v[8 * e + 4]), __uint_as_float(v[8 * e + 5]));
                    w[e].w = (int)pack_bf16x2(__uint_as_float(v[8 * e + 6]), __uint_as_float(v[8 * e + 7]));
                }
                if (p.dual_dact3) {
                    const int px = col >> 6, ch = col & 63;
                    const int oy = px / 7, ox = px - oy * 7;
                    int4* da = reinterpret_cast<int4*>(p.out + ((int64_t)r * 81 + oy * 9 + ox) * 64 + ch);
                    int4* db = reinterpret_cast<int4*>(p.out2 + ((int64_t)r * 121 + (oy + 2) * 11 + ox + 2) * 64 + ch);
                    da[0] = w[0]; da[1] = w[1]; da[2] = w[2]; da[3] = w[3];
                    db[0] = w[0]; db[1] = w[1]; db[2] = w[2]; db[3] = w[3];
                } else {
                    int4* dst = reinterpret_cast<int4*>(p.out + ooff + col);
                    dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ------------------------------------------------------------------ kernel 2d: TMA-fed weight gradient (fc)
// D[o, k] = sum_m dhid[m, o] * act3[m, k]: both operands are row-major, so each 64-row x 64-column chunk image is
// one TMA box; they are consumed as MN-major operands.  grid = (row splits, X groups of 2 chunks, Y groups of 4).
__global__ void __launch_bounds__(160, 1) tc_wgrad_tma(const __grid_constant__ CUtensorMap tmX,
                                                       const __grid_constant__ CUtensorMap tmY,
                                                       int64_t M, int64_t rows_per_cta, int nxc, int nyc, float* ws) {
    constexpr int R = 64, STAGES = 4;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NY = nyc * 64;
    const int xc0 = blockIdx.y * nxc, yc0 = blockIdx.z * nyc;
    const int xt = nxc / 2;
    constexpr int chunk_img = R * 128;
    const int stage_bytes = (nxc + nyc) * chunk_img;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < xt * NY) tmem_cols <<= 1;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&done_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmY);
    }
    if (warp == 4) tmem_alloc(&tmem_base_smem, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_smem;
    const int64_t m_begin = (int64_t)blockIdx.x * rows_per_cta;
    int64_t m_end = m_begin + rows_per_cta;
    if (m_end > M) m_end = M;
    const int nsteps = m_end > m_begin ? (int)((m_end - m_begin + R - 1) / R) : 0;

    if (warp == 0 && (tid & 31) == 0) {
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
            const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
            const int m0 = (int)(m_begin + (int64_t)it * R);
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            for (int c = 0; c < nxc; ++c) tma_load_2d(dst + c * chunk_img, &tmX, (xc0 + c) * 64, m0, &full_bar[s]);
            for (int c = 0; c < nyc; ++c) tma_load_2d(dst + (nxc + c) * chunk_img, &tmY, (yc0 + c) * 64, m0, &full_bar[s]);
        }
    } else if (warp == 4 && (tid & 31) == 0) {
        const uint32_t idesc = make_idesc(128, NY, 1, 1);
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            tc_fence_after_sync();
            const uint32_t xa = smem_u32(smem + (size_t)s * stage_bytes), ya = xa + nxc * chunk_img;
            for (int t = 0; t < xt; ++t) {
#pragma unroll
                for (int kk = 0; kk < R / 16; ++kk) {
                    const uint64_t adesc = desc_mnmajor(xa + (2 * t) * chunk_img + kk * 2048, chunk_img);
                    const uint64_t bdesc = desc_mnmajor(ya + kk * 2048, chunk_img);
                    umma_bf16(tmem_d + t * NY, adesc, bdesc, idesc, (it | kk) != 0);
                }
            }
            umma_commit(&empty_bar[s]);
        }
        umma_commit(&done_bar);
    }
    if (warp < 4) {
        if (nsteps > 0) {
            mbar_wait(&done_bar, 0);
            tc_fence_after_sync();
        }
        const int64_t KXtot = (int64_t)gridDim.y * nxc * 64, NYtot = (int64_t)gridDim.z * NY;
        float* wsb = ws + (int64_t)blockIdx.x * KXtot * NYtot;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int t = 0; t < xt; ++t) {
            float* dst = wsb + ((int64_t)xc0 * 64 + t * 128 + tid) * NYtot + (int64_t)yc0 * 64;
            for (int c0 = 0; c0 < NY; c0 += 16) {
                uint32_t v[16];
                if (nsteps > 0) {
                    tmem_ld16(lane_addr + t * NY + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0u;
                }
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    *reinterpret_cast<float4*>(dst + c0 + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                             __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_d, tmem_cols);
}

// ---- host: tensor maps for row-major bf16 matrices (cuTensorMapEncodeTiled resolved through the runtime)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int make_tmap_2d(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int box_rows, const char* what) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
            return fail(B200RL_ERR_CUDA, "%s: cuTensorMapEncodeTiled not available (%s)", what, cudaGetErrorString(e));
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200RL_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d)", what, (int)r);
    return B200RL_OK;
}

// [n_images][rows_per_image][cols] bf16, box = [1][box_rows][64]: rows past an image's end are zero-filled
static int make_tmap_3d(CUtensorMap* tm, const void* base, int64_t n_images, int64_t rows_per_image, int64_t cols, int box_rows,
                        const char* what) {
    if (!g_encode) {
        CUtensorMap dummy;
        int rc = make_tmap_2d(&dummy, base, 128, 64, 8, what);      // resolves the driver entry point
        if (rc) return rc;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_per_image, (cuuint64_t)n_images};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows_per_image * (cuuint64_t)cols * 2};
    const cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200RL_ERR_CUDA, "%s: cuTensorMapEncodeTiled (3-D) failed (%d)", what, (int)r);
    return B200RL_OK;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// A: row-major [M, 64*nchunks] bf16 (p.A), weights p.Bw [N, 64*nchunks]; epilogue fields as tc_gemm_ws
template <int BN, int STAGES>
static int launch_gemm_tma(const KGemmParams& p, cudaStream_t s, const char* what) {
    const size_t smem = (size_t)STAGES * (128 * 128 + BN * 128) + 1024;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_tma<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "%s: smem attribute (%zu B): %s", what, smem, cudaGetErrorString(e));
        attr = smem;
    }
    CUtensorMap tmA, tmB;
    int rc;
    const int64_t K = (int64_t)p.nchunks * 64;
    if ((rc = make_tmap_2d(&tmA, p.A, p.M, K, 128, what))) return rc;
    if ((rc = make_tmap_2d(&tmB, p.Bw, p.N, K, BN, what))) return rc;
    const int ntn = (int)ceil_div(p.N, BN);
    const int total = (int)ceil_div(p.M, 128) * ntn;
    int grid = num_sms();
    if (grid > total) grid = total;
    tc_gemm_tma<BN, STAGES><<<grid, 320, smem, s>>>(tmA, tmB, p, total, ntn);
    return check_launch(what);
}

template <int BN, int CPR, int STAGES, int NTAPS>
static int launch_conv_win(const WinParams& p, cudaStream_t s, const char* what) {
    if (p.ntaps != NTAPS) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: %d taps, kernel instance has %d", what, p.ntaps, NTAPS);
    const size_t smem = (size_t)p.ntaps * CPR * BN * 128 + (size_t)STAGES * p.WR * 128 * CPR + 1024;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv_win<BN, CPR, STAGES, NTAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "%s: smem attribute (%zu B): %s", what, smem, cudaGetErrorString(e));
        attr = smem;
    }
    if ((int64_t)p.G * p.Wp >= 65536 || p.G < 1 || p.N % 32 != 0)
        return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: grid %d x width %d outside the epilogue's multiply-shift range, or N %% 32 != 0", what, p.G, p.Wp);
    if (p.rows && !p.tpi_shift) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: the image gather needs image-aligned tiling", what);
    if (p.tpi_shift && (128 << p.tpi_shift) < p.G) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: tiles per image too small", what);
    const int total = p.tpi_shift ? (int)((int64_t)p.n << p.tpi_shift) : (int)ceil_div(p.M, 128);
    int grid = num_sms();
    if (grid > total) grid = total;
    CUtensorMap tmA;
    memset(&tmA, 0, sizeof(tmA));
    int rc;
    // the window is a TMA box [WR rows x 64 channels] per column chunk: of the linear grid, or of one image
    if (p.tpi_shift) rc = make_tmap_3d(&tmA, p.A, p.n_images, p.G, (int64_t)CPR * 64, p.WR, what);
    else rc = make_tmap_2d(&tmA, p.A, p.M, (int64_t)CPR * 64, p.WR, what);
    if (rc) return rc;
    tc_conv_win<BN, CPR, STAGES, NTAPS><<<grid, 320, smem, s>>>(tmA, p, total);
    return check_launch(what);
}

// ------------------------------------------------------------------ kernel 2c: window weight gradient
// dW^T[(tap,channel), co] = sum over grid rows r of X[r + shift_tap, channel] * dY[r, co] with X and dY on the
// SAME linear grid (dY is zero at positions that are not valid outputs).  Per step of 128 rows the CTA
// stages one X window (128 + max shift rows) and 128 dY rows; every tap is an MN-major descriptor shifted by
// whole rows.  Output tile t pairs the 64-channel chunks slot[2t], slot[2t+1].
constexpr int kWgradWinStages = 3;       // stage = X window (<= 37 KB) + 16 KB of dY rows; 3 stages keep conv1 at 2 CTAs per SM (4 were measured 40 % slower)
struct WGradWinParams {
    const bf16* X; const int64_t* rows; int64_t M; int n, G;
    int tpi_shift;           // > 0: image-aligned steps (2^tpi_shift steps of 128 rows per image, M = n << (7 + tpi_shift))
    int64_t n_images;        // images addressable through `rows`
    int cpr;                 // 64-channel column chunks per X row
    int nslots;              // even; chunk of slot s = (tap slot_tap[s], column chunk slot_cc[s])
    int slot_tap[16], slot_cc[16];
    int shift[16];           // per tap
    int WRX;                 // X window rows
    const bf16* Y; int ldy, ncolsY;
    int64_t rows_per_cta;    // multiple of 128
    float* ws;               // [gridDim.x][nslots*64][64]
    float* wsb;              // [gridDim.x][64] bias-gradient partials: sum_r dY[r, co]
};

__global__ void __launch_bounds__(192, 1) tc_wgrad_win(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                       const WGradWinParams p, int use_tma) {
    constexpr int R = 128, STAGES = kWgradWinStages, LOOKAHEAD = 1, NY = 64;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], done_bar;
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int IMGX = p.WRX * 128;
    const int XBYTES = IMGX * p.cpr;
    const int stage_bytes = XBYTES + R * 128;
    const int xt = p.nslots / 2;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < xt * NY) tmem_cols <<= 1;
    float* sRed = reinterpret_cast<float*>(smem + (size_t)STAGES * stage_bytes);     // [16][64] bias partials (4 KB)
    if (tid == 0) {
        // full:  one expect_tx arrival (TMA) [+ the four cp.async warps that stage dY in image-aligned mode]
        // empty: the MMA commit [+ the four dY-summing warps when they read the stage after the TMA landed]
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], use_tma ? 1 : 5); mbar_init(&empty_bar[s], use_tma ? 5 : 1); }
        mbar_init(&done_bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmX);
        if (use_tma) tma_prefetch_desc(&tmY);
    }
    if (warp == 4) tmem_alloc(&tmem_base_smem, tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_smem;
    const int64_t m_begin = (int64_t)blockIdx.x * p.rows_per_cta;
    int64_t m_end = m_begin + p.rows_per_cta;
    if (m_end > p.M) m_end = p.M;
    const int nsteps = m_end > m_begin ? (int)((m_end - m_begin + R - 1) / R) : 0;
    const int tmask = (1 << p.tpi_shift) - 1;
    const int64_t g0 = m_begin / R;                       // first global step of this CTA (image-aligned mode)

    if (warp == 5) {
        // ======================= TMA producer (one lane): X window [+ dY rows when they are 128 bytes wide] =========
        if ((tid & 31) == 0) {
            int z_next = 0;
            if (!use_tma && nsteps > 0) {
                const int64_t img = g0 >> p.tpi_shift;
                z_next = p.rows ? (int)__ldg(p.rows + (img < p.n ? img : 0)) : (int)img;
            }
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % STAGES;
                const int z = z_next;
                if (!use_tma && it + 1 < nsteps) {        // gather index of the next step, one step ahead
                    const int64_t img1 = (g0 + it + 1) >> p.tpi_shift;
                    z_next = p.rows ? (int)__ldg(p.rows + (img1 < p.n ? img1 : 0)) : (int)img1;
                }
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
                if (use_tma) {
                    const int m0 = (int)(m_begin + (int64_t)it * R);
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    for (int c = 0; c < p.cpr; ++c) tma_load_2d(dst + c * IMGX, &tmX, c * 64, m0, &full_bar[s]);
                    tma_load_2d(dst + XBYTES, &tmY, 0, m0, &full_bar[s]);
                } else {
                    const int t_in = (int)((g0 + it) & tmask);
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)XBYTES);
                    for (int c = 0; c < p.cpr; ++c) tma_load_3d(dst + c * IMGX, &tmX, c * 64, t_in * 128, z, &full_bar[s]);
                }
            }
        }
    } else if (warp < 4) {
        // ======================= dY warps: bias gradient = column sums of dY, taken from the staged tile ==========
        // Thread (tid>>3, tid&7) owns rows ps*16 + (tid>>3) and the 16-byte chunk (tid&7) = 8 channels of every step;
        // it adds them up in fp32 (fixed order).  This replaces an all-ones MMA per 16 rows, which cost a quarter to a
        // third of the kernel's shared-memory operand bandwidth.  In image-aligned mode (conv1: dY rows are 64 bytes,
        // no 128-byte TMA box) the same threads first copy those chunks in with cp.async.
        const int rq = tid >> 3, c16 = tid & 7;
        float bsum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) bsum[e] = 0.f;
        auto add_step = [&](const uint8_t* sYp) {
#pragma unroll
            for (int ps = 0; ps < R / 16; ++ps) {
                const int rr = ps * 16 + rq;
                const int4 v = *reinterpret_cast<const int4*>(sYp + img_off(rr, c16));
                const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bsum[2 * e] += __uint_as_float(w[e] << 16);
                    bsum[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
                }
            }
        };
        if (use_tma) {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % STAGES;
                mbar_wait(&full_bar[s], (it / STAGES) & 1);
                add_step(smem + (size_t)s * stage_bytes + XBYTES);
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
            }
        } else {
            for (int it = 0; it < nsteps; ++it) {
                const int s = it % STAGES;
                const int64_t g = g0 + it;
                const int64_t img = g >> p.tpi_shift;
                const int t_in = (int)(g & tmask);
                if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
                const uint32_t sY = smem_u32(smem + (size_t)s * stage_bytes + XBYTES);
                // dY rows of this step (zero past the image's G rows: those grid positions are padding)
#pragma unroll
                for (int ps = 0; ps < R / 16; ++ps) {
                    const int rr = ps * 16 + rq;
                    const int rl = t_in * 128 + rr;
                    const int col = c16 * 8;
                    const bool ok = rl < p.G && img < p.n && col < p.ncolsY;
                    cp_async16(sY + img_off(rr, c16), p.Y + (ok ? (img * p.G + rl) * (int64_t)p.ldy + col : 0), ok ? 16u : 0u);
                }
                cp_async_commit();
                if (it >= LOOKAHEAD) {
                    cp_async_wait<LOOKAHEAD>();
                    const int sd = (it - LOOKAHEAD) % STAGES;
                    add_step(smem + (size_t)sd * stage_bytes + XBYTES);      // this thread's own chunks have landed
                    fence_proxy_async_smem();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&full_bar[sd]);
                }
            }
            cp_async_wait<0>();
            for (int d = (nsteps >= LOOKAHEAD ? nsteps - LOOKAHEAD : 0); d < nsteps; ++d)
                add_step(smem + (size_t)(d % STAGES) * stage_bytes + XBYTES);
            fence_proxy_async_smem();
            __syncwarp();
            if ((tid & 31) == 0)
                for (int d = (nsteps >= LOOKAHEAD ? nsteps - LOOKAHEAD : 0); d < nsteps; ++d) mbar_arrive(&full_bar[d % STAGES]);
        }
        // fold the 16 row lanes of every column chunk in fixed order -> 64 bias partials of this CTA
#pragma unroll
        for (int e = 0; e < 8; ++e) sRed[rq * 64 + c16 * 8 + e] = bsum[e];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid < 64) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < 16; ++l) t += sRed[l * 64 + tid];
            p.wsb[(int64_t)blockIdx.x * NY + tid] = t;
        }
    } else {
        // ======================= MMA issuer: the whole warp walks the step loop (uniform control flow), one elected
        // lane issues.  Everything that does not depend on the stage is hoisted: per output tile the X operand's offset
        // inside the stage and its LBO field; descriptors then differ only in the 14-bit start-address field.
        const bool leader = elect_one();
        constexpr uint32_t idesc = make_idesc(128, NY, 1, 1);
        const uint64_t desc_hi = desc_mnmajor(0, 0) & 0xFFFFFFFF00000000ull;
        uint32_t arel[8], albo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            arel[t] = 0; albo[t] = 0;
            if (t < xt) {
                const uint32_t r0 = (uint32_t)(p.slot_cc[2 * t] * IMGX + p.shift[p.slot_tap[2 * t]] * 128);
                const uint32_t r1 = (uint32_t)(p.slot_cc[2 * t + 1] * IMGX + p.shift[p.slot_tap[2 * t + 1]] * 128);
                arel[t] = r0 >> 4;
                albo[t] = (((r1 - r0) >> 4) & 0x3FFFu) << 16;
            }
        }
        const uint32_t ylbo = (uint32_t)(((R * 128) >> 4) & 0x3FFF) << 16;
        for (int it = 0; it < nsteps; ++it) {
            const int s = it % STAGES;
            mbar_wait(&full_bar[s], (it / STAGES) & 1);
            tc_fence_after_sync();
            if (leader) {
                const uint32_t xa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t xa16 = (xa & 0x3FFFFu) >> 4, ya16 = (((xa + XBYTES) & 0x3FFFFu) >> 4) | ylbo;
                const uint32_t accum = it != 0 ? 1u : 0u;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (t < xt) {
                        const uint32_t a_lo = (xa16 + arel[t]) | albo[t];
#pragma unroll
                        for (int kk = 0; kk < R / 16; ++kk)
                            umma_bf16(tmem_d + t * NY, desc_hi | (uint64_t)(a_lo + kk * 128), desc_hi | (uint64_t)(ya16 + kk * 128),
                                      idesc, kk != 0 ? 1u : accum);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            __syncwarp();
        }
        if (leader) umma_commit(&done_bar);
        __syncwarp();
    }
    if (warp < 4) {
        if (nsteps > 0) {
            mbar_wait(&done_bar, 0);
            tc_fence_after_sync();
        }
        float* wsb = p.ws + (int64_t)blockIdx.x * (p.nslots * 64) * NY;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int t = 0; t < xt; ++t) {
            float* dst = wsb + (int64_t)(t * 128 + tid) * NY;
            for (int c0 = 0; c0 < NY; c0 += 16) {
                uint32_t v[16];
                if (nsteps > 0) {
                    tmem_ld16(lane_addr + t * NY + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0u;
                }
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    *reinterpret_cast<float4*>(dst + c0 + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                             __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_d, tmem_cols);
}

// out[idx] = sum_z part[z][idx], z < nslabs: block = 8 z-lanes x 32 consecutive outputs; every z-lane sums its
// slabs (z = lane, lane+8, ...) in ascending order, lane 0 then adds the 8 lane sums in order (deterministic).
__device__ __forceinline__ float zlane_sum(const float* __restrict__ part, int64_t slab, int nslabs, int64_t idx, bool valid,
                                           float* red /* [256] */) {
    const int zl = threadIdx.x >> 5, ol = threadIdx.x & 31;
    float s = 0.f;
    if (valid) {
        int z = zl;
        for (; z + 24 < nslabs; z += 32) {
            const float v0 = part[(int64_t)z * slab + idx], v1 = part[(int64_t)(z + 8) * slab + idx];
            const float v2 = part[(int64_t)(z + 16) * slab + idx], v3 = part[(int64_t)(z + 24) * slab + idx];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; z < nslabs; z += 8) s += part[(int64_t)z * slab + idx];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    float t = 0.f;
    if (zl == 0) {
#pragma unroll
        for (int l = 0; l < 8; ++l) t += red[l * 32 + ol];
    }
    return t;                                           // meaningful for zl == 0
}

// fold for the window weight gradients: ws[S][nslots*64][64] -> torch layout dst[co][c][ky][kx].
//   layer 1: slot = tap (a,b); row channel q = c*16 + sy*4 + sx; ky = 4a+sy, kx = 4b+sx; 32 outputs, Cin 4, 8x8
//   layer 2: slot = (tap (a,b), cc); q = cc*64 + row = (py*2+px)*32 + c; ky = 2a+py, kx = 2b+px; Cin 32, 4x4
//   layer 3: slot -> tap (ky,kx) via slot_tap (a duplicate slot is skipped); q = c; Cin 64, 3x3
struct FoldWin { int layer, S, nslots, Cout; int slot_tap[16], slot_cc[16], slot_skip[16]; float scale;
                 const float* wsb; float* db; };
__global__ void __launch_bounds__(256) tc_fold_win(const float* __restrict__ ws, const FoldWin f, float* __restrict__ dst) {
    __shared__ float red[256];
    const int idx = blockIdx.x * 32 + (threadIdx.x & 31);            // (slot*64 + row) * Cout + co
    const int KX = f.nslots * 64;
    if (idx >= KX * f.Cout) {                                        // trailing blocks fold the bias partials
        const int co = idx - KX * f.Cout;
        const bool valid = co < f.Cout && f.db != nullptr;
        const float s = zlane_sum(f.wsb, 64, f.S, co, valid, red);
        if (valid && threadIdx.x < 32) f.db[co] = s;
        return;
    }
    const int xi = idx / f.Cout, co = idx - xi * f.Cout;
    const int slot = xi >> 6, row = xi & 63;
    const bool valid = !f.slot_skip[slot];
    float s = zlane_sum(ws, (int64_t)KX * 64, f.S, (int64_t)xi * 64 + co, valid, red);
    if (!valid || threadIdx.x >= 32) return;
    s *= f.scale;
    const int tap = f.slot_tap[slot];
    int64_t o;
    if (f.layer == 1) {
        const int c = row >> 4, sy = (row >> 2) & 3, sx = row & 3;
        o = (((int64_t)co * 4 + c) * 8 + ((tap >> 1) * 4 + sy)) * 8 + (tap & 1) * 4 + sx;
    } else if (f.layer == 2) {
        const int q = f.slot_cc[slot] * 64 + row;
        const int g = q >> 5, c = q & 31;
        o = (((int64_t)co * 32 + c) * 4 + (2 * (tap >> 1) + (g >> 1))) * 4 + 2 * (tap & 1) + (g & 1);
    } else {
        o = ((int64_t)co * 64 + row) * 9 + tap;
    }
    dst[o] = s;
}

// fold the fc weight-gradient partials ws[S][KX rows = o][NY cols = k], k = p*64 + c, into the REFERENCE's layout
// dst[o][c*49 + p] (torch flattens NCHW activations channel-major); partial slabs are added in ascending order.
__global__ void tc_fold_fc(const float* __restrict__ ws, int S, int KX, int NY, int validX, int validY,
                           int C, int KK, float scale, float* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)validX * validY;
    if (idx >= total) return;
    const int xi = (int)(idx / validY), yi = (int)(idx - (int64_t)xi * validY);
    float s = 0.f;
    const int64_t slab = (int64_t)KX * NY;
    for (int z = 0; z < S; ++z) s += ws[z * slab + (int64_t)xi * NY + yi];
    const int pp = yi / C, c = yi - pp * C;
    dst[(int64_t)xi * validY + (int64_t)c * KK + pp] = s * scale;
}

// column sums of a bf16 matrix [M, ld] (bias gradients): two-level deterministic reduction.
// Block = 256 threads = (256 / (ncols/8)) row lanes x (ncols/8) column groups; every thread streams
// 16-byte vectors (8 columns) down its rows, then the row lanes are folded through shared memory.
__global__ void __launch_bounds__(256) tc_colsum_partial(const bf16* __restrict__ Y, int64_t M, int ld, int ncols,
                                                         int64_t rows_per_block, float* __restrict__ part) {
    __shared__ float red[256 * 8];
    const int cg = ncols >> 3;                 // column groups of 8
    const int lanes = 256 / cg;                // row lanes per block
    const int tx = threadIdx.x % cg, ty = threadIdx.x / cg;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    if (ty < lanes) {
        for (int64_t r = r0 + ty; r < r1; r += lanes) {
            const int4 v = ldg16(Y + r * ld + tx * 8);
            const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[2 * e] += __uint_as_float(w[e] << 16);
                acc[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = acc[e];
    __syncthreads();
    for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
        const int g = c >> 3, e = c & 7;
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(l * cg + g) * 8 + e];
        part[(int64_t)blockIdx.x * ncols + c] = s;
    }
}
__global__ void __launch_bounds__(256) tc_colsum_final(const float* __restrict__ part, int nblocks, int ncols,
                                                       float* __restrict__ db) {
    __shared__ float red[256];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const float s = zlane_sum(part, ncols, nblocks, c, c < ncols, red);
    if (c < ncols && threadIdx.x < 32) db[c] = s;
}

// uint8 frames [n,4,84,84] (NCHW, as envpool delivers them) -> space-to-depth bf16 [n,21,21,64] with
// channel = c*16 + sy*4 + sx for source pixel (4Y+sy, 4X+sx).  conv1 (8x8, stride 4) becomes a 2x2,
// stride-1 convolution over 64-channel NHWC pixels, i.e. the same 128-byte-per-tap gather as conv2/conv3.
// Done ONCE per environment step; the minibatch updates then read the bf16 rollout directly.
__global__ void __launch_bounds__(256) tc_frames_to_s2d(const uint8_t* __restrict__ obs, const int64_t* __restrict__ rows,
                                                        int64_t n, bf16* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // ((i*21 + Y)*21 + X)*4 + c
    if (idx >= n * 21 * 21 * 4) return;
    const int c = (int)(idx & 3);
    int64_t t = idx >> 2;
    const int X = (int)(t % 21); t /= 21;
    const int Y = (int)(t % 21);
    const int64_t i = t / 21;
    const int64_t img = rows ? rows[i] : i;
    const uint8_t* src = obs + img * 28224 + c * 7056 + (Y * 4) * 84 + X * 4;
    uint32_t o[8];
#pragma unroll
    for (int sy = 0; sy < 4; ++sy) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(src + sy * 84));
        o[2 * sy] = pack_bf16x2((float)(w & 0xFF), (float)((w >> 8) & 0xFF));
        o[2 * sy + 1] = pack_bf16x2((float)((w >> 16) & 0xFF), (float)(w >> 24));
    }
    int4* dst = reinterpret_cast<int4*>(out + (idx >> 2) * 64 + c * 16);
    dst[0] = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
    dst[1] = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
}

// ------------------------------------------------------------------ weight packing (fp32 master -> bf16 GEMM operands)
// conv weight w[co][c][ky][kx] -> fwd[co][(ky,kx,c)] (nhwc_k) or [co][(c,ky,kx)] (conv1), and
// dgrad[c][(ky,kx,co)] with taps FLIPPED implicitly by the loader's negative offsets (no flip needed here).
__global__ void tc_pack_conv(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int nchw_k,
                             bf16* __restrict__ fwd, bf16* __restrict__ dgrad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)Cout * Cin * KH * KW;
    if (idx >= total) return;
    int kx = (int)(idx % KW); int64_t t = idx / KW;
    int ky = (int)(t % KH); t /= KH;
    int c = (int)(t % Cin); int co = (int)(t / Cin);
    const bf16 v = __float2bfloat16(w[idx]);
    const int K = Cin * KH * KW;
    if (nchw_k) fwd[(int64_t)co * K + (c * KH + ky) * KW + kx] = v;
    else fwd[(int64_t)co * K + (ky * KW + kx) * Cin + c] = v;
    if (dgrad) dgrad[(int64_t)c * (KH * KW * Cout) + (ky * KW + kx) * Cout + co] = v;
}
// conv1 weight w[co][c][ky][kx] (8x8) -> [co][(a,b), c*16 + sy*4 + sx] with ky = 4a+sy, kx = 4b+sx
// (K order of the space-to-depth frames)
__global__ void tc_pack_conv1_s2d(const float* __restrict__ w, bf16* __restrict__ fwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * 256) return;
    const int kx = idx & 7, ky = (idx >> 3) & 7, c = (idx >> 6) & 3, co = idx >> 8;
    const int a = ky >> 2, sy = ky & 3, b = kx >> 2, sx = kx & 3;
    fwd[co * 256 + (a * 2 + b) * 64 + c * 16 + sy * 4 + sx] = __float2bfloat16(w[idx]);
}
// conv2 weight w[co][c][ky][kx] (4x4, stride 2) -> [co][(a,b) tap][(py,px,c)] with ky = 2a+py, kx = 2b+px:
// K order of the 2x2-cell (space-to-depth 2) activations [n,10,10,128]
__global__ void tc_pack_conv2_cells(const float* __restrict__ w, bf16* __restrict__ fwd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * 512) return;
    const int kx = idx & 3, ky = (idx >> 2) & 3, c = (idx >> 4) & 31, co = idx >> 9;
    const int a = ky >> 1, py = ky & 1, b = kx >> 1, px = kx & 1;
    fwd[co * 512 + (a * 2 + b) * 128 + (py * 2 + px) * 32 + c] = __float2bfloat16(w[idx]);
}
// conv2 data-gradient weights per stride-parity class: dg[cls][c][(a,b,co)] = w[co][c][py+2a][px+2b]
__global__ void tc_pack_conv_s2_classes(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ dg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)4 * Cin * 4 * Cout;
    if (idx >= total) return;
    int co = (int)(idx % Cout); int64_t t = idx / Cout;
    int ab = (int)(t % 4); t /= 4;
    int c = (int)(t % Cin); int cls = (int)(t / Cin);
    const int py = cls >> 1, px = cls & 1, a = ab >> 1, b = ab & 1;
    dg[idx] = __float2bfloat16(w[(((int64_t)co * Cin + c) * 4 + (py + 2 * a)) * 4 + (px + 2 * b)]);
}
// fc weight w[o][c*49+p] -> fwd[o][p*64+c]; dgrad[p*64+c][o].  Block = 8 output rows x 7 pixels x all 64 channels,
// staged through shared memory so that both packed layouts are written with 16-byte stores
// (fwd: 8 consecutive c of one (o, p); dgrad: the 8 o of one (p, c)).  Requires O % 8 == 0, PP % 7 == 0, C == 64.
__global__ void __launch_bounds__(256) tc_pack_fc(const float* __restrict__ w, int O, int PP, bf16* __restrict__ fwd,
                                                  bf16* __restrict__ dgrad) {
    constexpr int C = 64, PS = 7, R = 8;
    __shared__ float sw[R * C * PS];                    // [r][c][pl]
    const int K = C * PP;
    const int o0 = blockIdx.x * R, p0 = blockIdx.y * PS;
    for (int i = threadIdx.x; i < R * C * PS; i += blockDim.x) {
        const int pl = i % PS, rc = i / PS;             // rc = r*64 + c
        sw[i] = w[(int64_t)(o0 + (rc >> 6)) * K + (rc & 63) * PP + p0 + pl];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * PS * (C / 8); i += blockDim.x) {      // fwd[o][p*64 + c0 .. c0+7]
        const int g = i & 7, pl = (i >> 3) % PS, r = i / (8 * PS);
        const float* src = sw + (r * C + g * 8) * PS + pl;
        int4 v;
        v.x = (int)pack_bf16x2(src[0 * PS], src[1 * PS]); v.y = (int)pack_bf16x2(src[2 * PS], src[3 * PS]);
        v.z = (int)pack_bf16x2(src[4 * PS], src[5 * PS]); v.w = (int)pack_bf16x2(src[6 * PS], src[7 * PS]);
        *reinterpret_cast<int4*>(fwd + (int64_t)(o0 + r) * K + (p0 + pl) * C + g * 8) = v;
    }
    for (int i = threadIdx.x; i < PS * C; i += blockDim.x) {                 // dgrad[p*64 + c][o0 .. o0+7]
        const int c = i & 63, pl = i >> 6;
        const float* src = sw + c * PS + pl;
        int4 v;
        v.x = (int)pack_bf16x2(src[0 * C * PS], src[1 * C * PS]); v.y = (int)pack_bf16x2(src[2 * C * PS], src[3 * C * PS]);
        v.z = (int)pack_bf16x2(src[4 * C * PS], src[5 * C * PS]); v.w = (int)pack_bf16x2(src[6 * C * PS], src[7 * C * PS]);
        *reinterpret_cast<int4*>(dgrad + ((int64_t)(p0 + pl) * C + c) * O + o0) = v;
    }
}

// ------------------------------------------------------------------ policy/value heads (tiny: CUDA cores, fp32 math)
// A1 = A + 1 head outputs (logits | value), 1 <= A1 <= kMaxHeads.  Head weights live in dynamic shared memory.
constexpr int kMaxHeads = 32;
constexpr int kHeadsPartialBlocks = 296;     // row blocks of the head weight gradient (x2 row lanes = partial slabs)
// rows per block of tc_heads_bwd_weight: its dhead rows are staged in (static-limit) shared memory, <= 256 x 32 floats
static inline int64_t heads_rows_per_block(int64_t n) {
    int64_t rpb = (n + kHeadsPartialBlocks - 1) / kHeadsPartialBlocks;
    if (rpb < 16) rpb = 16;
    if (rpb > 256) rpb = 256;
    return rpb;
}

// out[n][A1] = hidden[n][512](bf16) . Wh[A1][512]^T + bh.  One warp per row: lane l holds hidden units
// [8l, 8l+8) and [256+8l, 256+8l+8) (two 16-byte loads), weights are read as float4 from shared memory.
__global__ void __launch_bounds__(256) tc_heads_fwd(const bf16* __restrict__ hid, const float* __restrict__ Wh,
                                                    const float* __restrict__ bh, int64_t n, int A1, int H,
                                                    float* __restrict__ out) {
    extern __shared__ float sW[];                       // [A1][512]
    for (int i = threadIdx.x; i < A1 * 512; i += blockDim.x) sW[i] = Wh[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * wpb) {
        float hv[16];
        const bf16* hp = hid + row * 512;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int4 v = ldg16(hp + q * 256 + lane * 8);
            const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hv[q * 8 + 2 * e] = __uint_as_float(w[e] << 16);
                hv[q * 8 + 2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
            }
        }
        float mine = 0.f;                               // lane a keeps output a
        for (int a = 0; a < A1; ++a) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 w0 = *reinterpret_cast<const float4*>(sW + a * 512 + q * 256 + lane * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sW + a * 512 + q * 256 + lane * 8 + 4);
                s = fmaf(hv[q * 8 + 0], w0.x, s); s = fmaf(hv[q * 8 + 1], w0.y, s);
                s = fmaf(hv[q * 8 + 2], w0.z, s); s = fmaf(hv[q * 8 + 3], w0.w, s);
                s = fmaf(hv[q * 8 + 4], w1.x, s); s = fmaf(hv[q * 8 + 5], w1.y, s);
                s = fmaf(hv[q * 8 + 6], w1.z, s); s = fmaf(hv[q * 8 + 7], w1.w, s);
            }
            s = warp_sum(s);
            if (lane == a) mine = s + bh[a];
        }
        if (lane < A1) out[row * A1 + lane] = mine;     // one coalesced store per row
    }
}
// dhid_pre[n][512] (bf16) = (dhead[n][A1] . Wh[A1][512]) * (hid > 0).  Thread = 8 consecutive hidden units of one
// row (one mask byte in, one 16-byte store out).  Weights are staged transposed, sWt[a][e][group], so the 32 lanes
// of a warp (consecutive groups) hit 32 different banks.
__global__ void __launch_bounds__(256) tc_heads_bwd_data(const float* __restrict__ dhead, const float* __restrict__ Wh,
                                                         const uint8_t* __restrict__ hid_bits, int64_t n, int A1, int H,
                                                         bf16* __restrict__ dhid) {
    extern __shared__ float sWt[];                      // [A1][8][64]
    for (int i = threadIdx.x; i < A1 * 512; i += blockDim.x) {
        const int a = i >> 9, h = i & 511;
        sWt[a * 512 + (h & 7) * 64 + (h >> 3)] = Wh[i];
    }
    __syncthreads();
    const int64_t total = n * 64;                      // 64 groups of 8 per row
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx >> 6;
        const int g = (int)(idx & 63);
        const uint32_t m = hid_bits[idx];               // bit e: hid[row][8g + e] > 0
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
        for (int a = 0; a < A1; ++a) {
            const float d = __ldg(dhead + row * A1 + a);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(d, sWt[a * 512 + e * 64 + g], o[e]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) if (!((m >> e) & 1u)) o[e] = 0.f;
        int4 w;
        w.x = (int)pack_bf16x2(o[0], o[1]); w.y = (int)pack_bf16x2(o[2], o[3]);
        w.z = (int)pack_bf16x2(o[4], o[5]); w.w = (int)pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<int4*>(dhid + row * 512 + g * 8) = w;
    }
}
// dWh[a][h] = sum_m dhead[m][a] * hid[m][h]; dbh[a] = sum_m dhead[m][a]  (partial slabs per (row block, row lane),
// folded by tc_heads_fold).  Block = 512 threads = 2 row lanes x 256 hidden pairs; the block's dhead rows are staged
// in shared memory once, 8 rows of hidden values are in flight per thread.
template <int MAXA>
__global__ void __launch_bounds__(512) tc_heads_bwd_weight(const float* __restrict__ dhead, const bf16* __restrict__ hid,
                                                           int64_t n, int A1, int H, int64_t rows_per_block,
                                                           float* __restrict__ part) {
    extern __shared__ float sD[];                       // [rows_per_block][A1]
    const int hp = threadIdx.x & 255, rl = threadIdx.x >> 8;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > n) r1 = n;
    const int nrows = (int)(r1 > r0 ? r1 - r0 : 0);
    for (int i = threadIdx.x; i < nrows * A1; i += blockDim.x) sD[i] = dhead[r0 * A1 + i];
    __syncthreads();
    float acc0[MAXA], acc1[MAXA], bacc[MAXA];
#pragma unroll
    for (int a = 0; a < MAXA; ++a) { acc0[a] = 0.f; acc1[a] = 0.f; bacc[a] = 0.f; }
    const uint32_t* h2 = reinterpret_cast<const uint32_t*>(hid);      // bf16 pairs
    int r = rl;
    for (; r + 14 < nrows; r += 16) {                   // 8 rows (stride 2) in flight
        uint32_t hv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) hv[u] = __ldg(h2 + (r0 + r + 2 * u) * 256 + hp);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float x0 = __uint_as_float(hv[u] << 16), x1 = __uint_as_float(hv[u] & 0xFFFF0000u);
            const float* dr = sD + (r + 2 * u) * A1;
#pragma unroll
            for (int a = 0; a < MAXA; ++a) {
                if (a < A1) {
                    const float d = dr[a];
                    acc0[a] = fmaf(d, x0, acc0[a]); acc1[a] = fmaf(d, x1, acc1[a]); bacc[a] += d;
                }
            }
        }
    }
    for (; r < nrows; r += 2) {
        const uint32_t hv = __ldg(h2 + (r0 + r) * 256 + hp);
        const float x0 = __uint_as_float(hv << 16), x1 = __uint_as_float(hv & 0xFFFF0000u);
        const float* dr = sD + r * A1;
#pragma unroll
        for (int a = 0; a < MAXA; ++a) {
            if (a < A1) {
                const float d = dr[a];
                acc0[a] = fmaf(d, x0, acc0[a]); acc1[a] = fmaf(d, x1, acc1[a]); bacc[a] += d;
            }
        }
    }
    float* pb = part + ((int64_t)blockIdx.x * 2 + rl) * A1 * (H + 2);      // slab rows: H weights, bias, pad
#pragma unroll
    for (int a = 0; a < MAXA; ++a) {
        if (a < A1) {
            *reinterpret_cast<float2*>(pb + (int64_t)a * (H + 2) + 2 * hp) = make_float2(acc0[a], acc1[a]);
            if (hp == 0) pb[(int64_t)a * (H + 2) + H] = bacc[a];
        }
    }
}
__global__ void __launch_bounds__(256) tc_heads_fold(const float* __restrict__ part, int nslabs, int A1, int H,
                                                     float* __restrict__ dW, float* __restrict__ db) {
    __shared__ float red[256];
    const int idx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int a = idx / (H + 2), h = idx - a * (H + 2);
    const bool valid = a < A1 && h <= H;                // h == H: bias; h == H + 1: padding
    const float s = zlane_sum(part, (int64_t)A1 * (H + 2), nslabs, idx, valid, red);
    if (!valid || threadIdx.x >= 32) return;
    if (h == H) db[a] = s; else dW[(int64_t)a * H + h] = s;
}

}  // namespace b200rl

// =====================================================================================
// Host side: NatureCNN plan over the kernels above (C-ABI entry points, include/b200rl.h)
// =====================================================================================
namespace b200rl {

struct NatureLayout {
    int A;
    // flat fp32 parameter offsets (libb200rl order: trunk, then both head weights, then both head biases)
    int64_t c1w, c1b, c2w, c2b, c3w, c3b, fcw, fcb, hw, hb, total;
    // packed bf16 operand offsets (elements)
    int64_t w1f, w2f, w2dg, w3f, w3dg, wfcf, wfcdg, packed_total;
    explicit NatureLayout(int A_) : A(A_) {
        int64_t o = 0;
        c1w = o; o += 32 * 4 * 8 * 8;  c1b = o; o += 32;
        c2w = o; o += 64 * 32 * 4 * 4; c2b = o; o += 64;
        c3w = o; o += 64 * 64 * 3 * 3; c3b = o; o += 64;
        fcw = o; o += 512 * 3136;      fcb = o; o += 512;
        hw = o;  o += (int64_t)(A + 1) * 512;
        hb = o;  o += A + 1;
        total = o;
        int64_t q = 0;
        w1f = q; q += 32 * 256;
        w2f = q; q += 64 * 512;
        w2dg = q; q += 4 * 32 * 256;
        w3f = q; q += 64 * 576;
        w3dg = q; q += 64 * 576;
        wfcf = q; q += 512 * 3136;
        wfcdg = q; q += 3136 * 512;
        packed_total = q;
    }
};

struct NatureActs {   // bf16 element offsets inside the (zero-initialised) activation workspace for batch n
    int64_t x0, act1, act2, act3, hid, dhid, dact3a, dact3b, dact2a, dact2b, dact1, m1, m2, m3, m4, total;
    explicit NatureActs(int64_t n, bool with_x0 = true) {
        int64_t o = 0;
        x0 = o; if (with_x0) o += n * 28224;     // space-to-depth frames [n,441,64] (only for uint8 input)
        act1 = o; o += n * 12800;                // conv1 out as 2x2 cells   [n,100,128]
        act2 = o; o += n * 5184;                 // conv2 out               [n, 81, 64]
        act3 = o; o += n * 3136;                 // conv3 out               [n, 49, 64]
        hid = o;  o += n * 512;
        dhid = o; o += n * 512;
        dact3a = o; o += n * 5184;               // d(act3) on the 9x9 linear grid (zeros outside 7x7)
        dact3b = o; o += n * 7744;               // d(act3) zero-padded to 11x11 (interior at +2,+2)
        dact2a = o; o += n * 6400;               // d(act2) on the 10x10 linear grid (zeros at row/col 9)
        dact2b = o; o += n * 7744;               // d(act2) zero-padded to 11x11 (interior at +1,+1)
        dact1 = o; o += n * 14112;               // d(act1) on the 21x21 linear grid, 32 channels
        // ReLU masks as bits (uint32 words; offsets stay in bf16 elements = 2 words per 4 elements)
        auto pad8 = [](int64_t v) { return (v + 7) & ~int64_t(7); };
        m1 = o; o += pad8(n * 100 * 4 * 2);      // act1 > 0: [n,100 cells] x 4 words (128 channels)
        m2 = o; o += pad8(n * 81 * 2 * 2);       // act2 > 0: [n,81] x 2 words
        m3 = o; o += pad8(n * 49 * 2 * 2);       // act3 > 0: [n,49] x 2 words (= dense [n,3136] / 32)
        m4 = o; o += pad8(n * 16 * 2);           // hid  > 0: [n] x 16 words
        total = o;
    }
};

static void gemm_rowmajor(KGemmParams& p, const bf16* x, int64_t n, int nchunks) {   // x [n, 64*nchunks]
    memset(&p, 0, sizeof(p));
    p.scale = 1.f;
    p.A = x; p.M = n; p.nchunks = nchunks;
}


// ---- window-convolution descriptions of the three conv layers
static void win_defaults(WinParams& p) { memset(&p, 0, sizeof(p)); p.scale = 1.f; }
static int round8(int v) { return (v + 7) & ~7; }
static void win_conv1(WinParams& p, const bf16* x0, const int64_t* rows, int64_t n) {     // 2x2 taps on the 21x21 s2d grid
    p.A = x0; p.rows = rows; p.n = (int)n; p.G = 441; p.Wp = 21; p.M = n * 441;
    p.tpi_shift = 2;                             // 4 tiles of 128 grid rows per image (441 used)
    p.n_images = rows ? (int64_t)1 << 24 : n;    // gather indices are the caller's contract (never range-checked)
    p.ntaps = 4; p.shift[0] = 0; p.shift[1] = 1; p.shift[2] = 21; p.shift[3] = 22; p.WR = round8(128 + 22);
}
static void win_conv2(WinParams& p, const bf16* act1, int64_t n) {                         // 2x2 taps on the 10x10 cell grid
    p.A = act1; p.n = (int)n; p.G = 100; p.Wp = 10; p.M = n * 100;
    p.ntaps = 4; p.shift[0] = 0; p.shift[1] = 1; p.shift[2] = 10; p.shift[3] = 11; p.WR = round8(128 + 11);
}
static void win_conv3(WinParams& p, const bf16* act2, int64_t n) {                         // 3x3 taps on the 9x9 grid
    p.A = act2; p.n = (int)n; p.G = 81; p.Wp = 9; p.M = n * 81;
    p.ntaps = 9; for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) p.shift[ky * 3 + kx] = ky * 9 + kx;
    p.WR = round8(128 + 20);
}
static void wgw_defaults(WGradWinParams& w) { memset(&w, 0, sizeof(w)); }

static int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

struct WPlan { int64_t rows_per_cta; int splits; };
static WPlan wgrad_plan(int64_t M, int target_ctas, int quantum = 32) {
    WPlan w;
    w.rows_per_cta = round_up(ceil_div(M, target_ctas), quantum);
    if (w.rows_per_cta < quantum) w.rows_per_cta = quantum;
    w.splits = (int)ceil_div(M, w.rows_per_cta);
    if (w.splits < 1) w.splits = 1;
    return w;
}
static const int kC1Ctas = 296, kC2Ctas = 148, kC3Ctas = 148, kFcSplits = 8;

static int launch_wgrad_win(const WGradWinParams& p, int ctas, cudaStream_t s, const char* what) {
    const size_t smem = (size_t)kWgradWinStages * ((size_t)p.WRX * 128 * p.cpr + 128 * 128) + 4096 + 1024;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad_win, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "%s: smem attribute (%zu B): %s", what, smem, cudaGetErrorString(e));
        attr = smem;
    }
    CUtensorMap tmX, tmY;
    memset(&tmX, 0, sizeof(tmX)); memset(&tmY, 0, sizeof(tmY));
    // TMA when rows are contiguous, dY rows are exactly 128 bytes and every CTA owns whole 128-row steps
    // all-TMA when dY rows are exactly 128 bytes; image-aligned steps (3-D TMA for X, cp.async for dY) otherwise
    const int use_tma = p.tpi_shift ? 0 : 1;
    if (p.rows_per_cta % 128 != 0) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: rows per CTA must be a multiple of 128", what);
    int rc;
    if (use_tma) {
        if (p.rows || p.ldy != 64 || p.ncolsY != 64)
            return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: linear-grid mode needs contiguous rows and 64-channel dY", what);
        if ((rc = make_tmap_2d(&tmX, p.X, p.M, (int64_t)p.cpr * 64, p.WRX, what))) return rc;
        if ((rc = make_tmap_2d(&tmY, p.Y, p.M, 64, 128, what))) return rc;
    } else {
        if ((128 << p.tpi_shift) < p.G) return fail(B200RL_ERR_INVALID_ARGUMENT, "%s: steps per image too small", what);
        if ((rc = make_tmap_3d(&tmX, p.X, p.n_images, p.G, (int64_t)p.cpr * 64, p.WRX, what))) return rc;
    }
    tc_wgrad_win<<<ctas, 192, smem, s>>>(tmX, tmY, p, use_tma);
    return check_launch(what);
}

static int colsum(const bf16* Y, int64_t M, int ld, int ncols, float* part, float* db, cudaStream_t s) {
    int64_t rpb = ceil_div(M, 148 * 3);
    if (rpb < 64) rpb = 64;
    const int nb = (int)ceil_div(M, rpb);
    tc_colsum_partial<<<nb, 256, 0, s>>>(Y, M, ld, ncols, rpb, part);
    tc_colsum_final<<<(unsigned)ceil_div(ncols, 32), 256, 0, s>>>(part, nb, ncols, db);
    return check_launch("colsum", 2);
}
static size_t colsum_ws(int64_t M, int ncols) {
    int64_t rpb = ceil_div(M, 148 * 3);
    if (rpb < 64) rpb = 64;
    return (size_t)ceil_div(M, rpb) * ncols * sizeof(float);
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int64_t b200rl_naturecnn_param_count(int A) { return A >= 1 ? NatureLayout(A).total : -1; }
extern "C" size_t b200rl_naturecnn_bf16_packed_bytes(int A) { return A >= 1 ? (size_t)NatureLayout(A).packed_total * 2 : 0; }
extern "C" size_t b200rl_naturecnn_bf16_acts_bytes(int64_t n, int obs_format) {
    return n >= 0 ? (size_t)NatureActs(n, obs_format == B200RL_OBS_U8_NCHW).total * 2 + 256 : 0;
}

extern "C" int b200rl_frames_to_s2d_bf16(const uint8_t* obs, const int64_t* rows, int64_t n, void* out, void* stream) {
    B200RL_REQUIRE(n >= 0, "frames_to_s2d: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(obs && out, "frames_to_s2d: null pointer");
    B200RL_REQUIRE(aligned(obs, 4) && aligned(out, 16), "frames_to_s2d: misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(s, "frames_to_s2d", 0, (double)n * 28224 * 3);
    tc_frames_to_s2d<<<(unsigned)ceil_div(n * 1764, 256), 256, 0, s>>>(obs, rows, n, reinterpret_cast<bf16*>(out));
    return check_launch("frames_to_s2d");
}

extern "C" size_t b200rl_naturecnn_bf16_workspace_bytes(int64_t n, int A) {
    if (n < 1 || A < 1) return 0;
    size_t a = 0;
    auto mx = [&](size_t v) { if (v > a) a = v; };
    mx((size_t)wgrad_plan(n * 512, kC1Ctas, 128).splits * 256 * 64 * 4);
    mx((size_t)wgrad_plan(n * 100, kC2Ctas, 128).splits * 512 * 64 * 4);
    mx((size_t)wgrad_plan(n * 81, kC3Ctas, 128).splits * 640 * 64 * 4);
    mx((size_t)wgrad_plan(n, kFcSplits, 64).splits * 512 * (13 * 256) * 4);
    size_t b = 0;
    auto mb = [&](size_t v) { if (v > b) b = v; };
    mb(colsum_ws(n * 441, 32)); mb(colsum_ws(n * 100, 64)); mb(colsum_ws(n * 81, 64)); mb(colsum_ws(n, 512));
    mb((size_t)2 * ceil_div(n, heads_rows_per_block(n)) * (A + 1) * 514 * 4);
    return a + b + 512;
}

extern "C" int b200rl_naturecnn_bf16_pack(const float* params, int A, void* packed, void* stream) {
    B200RL_REQUIRE(params && packed && A >= 1 && A < kMaxHeads, "naturecnn_pack: bad arguments (A must be in [1,31])");
    B200RL_REQUIRE(aligned(packed, 16), "naturecnn_pack: packed buffer must be 16-B aligned");
    const NatureLayout L(A);
    bf16* P = reinterpret_cast<bf16*>(packed);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(s, "pack_weights", 0, (double)L.total * 4 + (double)L.packed_total * 2);
    tc_pack_conv1_s2d<<<32, 256, 0, s>>>(params + L.c1w, P + L.w1f);
    tc_pack_conv2_cells<<<128, 256, 0, s>>>(params + L.c2w, P + L.w2f);
    tc_pack_conv_s2_classes<<<(unsigned)ceil_div(32768, 256), 256, 0, s>>>(params + L.c2w, 64, 32, P + L.w2dg);
    tc_pack_conv<<<(unsigned)ceil_div(36864, 256), 256, 0, s>>>(params + L.c3w, 64, 64, 3, 3, 0, P + L.w3f, P + L.w3dg);
    tc_pack_fc<<<dim3(512 / 8, 49 / 7), 256, 0, s>>>(params + L.fcw, 512, 49, P + L.wfcf, P + L.wfcdg);
    return check_launch("naturecnn_pack", 5);
}

extern "C" int b200rl_naturecnn_bf16_forward(const void* obs, int obs_format, const int64_t* rows, int64_t n, int A,
                                             const float* params, const void* packed, void* acts,
                                             float* head_out, void* stream) {
    B200RL_REQUIRE(n >= 0, "naturecnn_forward: negative n");
    if (n == 0) return B200RL_OK;
    B200RL_REQUIRE(obs && params && packed && acts && head_out, "naturecnn_forward: null pointer");
    B200RL_REQUIRE(A >= 1 && A < kMaxHeads, "naturecnn_forward: A=%d outside [1,31]", A);
    B200RL_REQUIRE(obs_format == B200RL_OBS_U8_NCHW || obs_format == B200RL_OBS_S2D_BF16, "naturecnn_forward: bad obs_format %d", obs_format);
    B200RL_REQUIRE(aligned(obs, 16) && aligned(acts, 16) && aligned(packed, 16), "naturecnn_forward: misaligned buffer");
    B200RL_REQUIRE(n <= (int64_t)1 << 22, "naturecnn_forward: n too large");
    const NatureLayout L(A);
    const NatureActs Q(n, obs_format == B200RL_OBS_U8_NCHW);
    const bf16* P = reinterpret_cast<const bf16*>(packed);
    bf16* act = reinterpret_cast<bf16*>(acts);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    KGemmParams p;
    WinParams wp;
    // conv1: 2x2 window conv on space-to-depth frames -> act1 as 2x2 cells [n,10,10,128]
    const bf16* x0 = reinterpret_cast<const bf16*>(obs);
    const int64_t* x0rows = rows;
    if (obs_format == B200RL_OBS_U8_NCHW) {
        if ((rc = b200rl_frames_to_s2d_bf16(reinterpret_cast<const uint8_t*>(obs), rows, n, act + Q.x0, stream))) return rc;
        x0 = act + Q.x0; x0rows = nullptr;
    }
    win_defaults(wp); win_conv1(wp, x0, x0rows, n);
    wp.Bw = P + L.w1f; wp.N = 32; wp.vH = 20; wp.vW = 20; wp.out_mode = WOUT_S2D2; wp.out = act + Q.act1;
    wp.bias = params + L.c1b; wp.scale = 1.0f / 255.0f; wp.relu = 1; wp.mask_out = reinterpret_cast<uint32_t*>(act + Q.m1);
    { ProfScope ps(s, "conv1_fwd", 2.0 * n * 400 * 32 * 256, (double)n * ((28224 + 12800) * 2 + 1600));
      if ((rc = launch_conv_win<32, 1, 9, 4>(wp, s, "naturecnn/conv1"))) return rc; }
    // conv2: 2x2 window conv on the 128-channel cells -> act2 [n,9,9,64]
    win_defaults(wp); win_conv2(wp, act + Q.act1, n);
    wp.Bw = P + L.w2f; wp.N = 64; wp.vH = 9; wp.vW = 9; wp.out_mode = WOUT_DENSE; wp.out = act + Q.act2;
    wp.bias = params + L.c2b; wp.relu = 1; wp.mask_out = reinterpret_cast<uint32_t*>(act + Q.m2);
    { ProfScope ps(s, "conv2_fwd", 2.0 * n * 81 * 64 * 512, (double)n * ((12800 + 5184) * 2 + 648));
      if ((rc = launch_conv_win<64, 2, 4, 4>(wp, s, "naturecnn/conv2"))) return rc; }
    // conv3: 3x3 window conv -> act3 [n,7,7,64]
    win_defaults(wp); win_conv3(wp, act + Q.act2, n);
    wp.Bw = P + L.w3f; wp.N = 64; wp.vH = 7; wp.vW = 7; wp.out_mode = WOUT_DENSE; wp.out = act + Q.act3;
    wp.bias = params + L.c3b; wp.relu = 1; wp.mask_out = reinterpret_cast<uint32_t*>(act + Q.m3);
    { ProfScope ps(s, "conv3_fwd", 2.0 * n * 49 * 64 * 576, (double)n * ((5184 + 3136) * 2 + 392));
      if ((rc = launch_conv_win<64, 1, 6, 9>(wp, s, "naturecnn/conv3"))) return rc; }
    // fc -> hidden [n,512]
    gemm_rowmajor(p, act + Q.act3, n, 49);
    p.Bw = P + L.wfcf; p.N = 512; p.out = act + Q.hid; p.ldo = 512; p.bias = params + L.fcb; p.relu = 1;
    p.mask_out = reinterpret_cast<uint32_t*>(act + Q.m4);
    { ProfScope ps(s, "fc_fwd", 2.0 * n * 512 * 3136, (double)n * (3136 + 512) * 2 + 512.0 * 3136 * 2);
      // small batches (rollout step): narrower N tiles => 4x more CTAs for the same work
      if (n <= 8192) { if ((rc = launch_gemm_tma<64, 8>(p, s, "naturecnn/fc"))) return rc; }
      else if ((rc = launch_gemm_tma<256, 4>(p, s, "naturecnn/fc"))) return rc; }
    // heads (fp32 math on CUDA cores): head_out [n, A+1] = [logits | value]
    { ProfScope ps(s, "heads_fwd", 2.0 * n * 512 * (A + 1), (double)n * (1024 + 4 * (A + 1)));
      int hb = (int)ceil_div(n, 8); if (hb > num_sms() * 8) hb = num_sms() * 8;
      tc_heads_fwd<<<hb, 256, (size_t)(A + 1) * 2048, s>>>(act + Q.hid, params + L.hw, params + L.hb, n, A + 1, 512, head_out); }
    return check_launch("naturecnn/heads");
}

extern "C" int b200rl_naturecnn_bf16_backward(const void* obs, int obs_format, const int64_t* rows, int64_t n, int A,
                                              const float* params, const void* packed, void* acts,
                                              const float* dhead, float* grads,
                                              void* workspace, size_t workspace_bytes, void* stream) {
    B200RL_REQUIRE(n >= 1, "naturecnn_backward: n must be >= 1");
    B200RL_REQUIRE(obs && params && packed && acts && dhead && grads && workspace, "naturecnn_backward: null pointer");
    B200RL_REQUIRE(A >= 1 && A < kMaxHeads, "naturecnn_backward: A=%d outside [1,31]", A);
    B200RL_REQUIRE(aligned(workspace, 16), "naturecnn_backward: workspace misaligned");
    const size_t need = b200rl_naturecnn_bf16_workspace_bytes(n, A);
    if (workspace_bytes < need) return fail(B200RL_ERR_WORKSPACE, "naturecnn_backward: workspace %zu < %zu", workspace_bytes, need);
    B200RL_REQUIRE(obs_format == B200RL_OBS_U8_NCHW || obs_format == B200RL_OBS_S2D_BF16, "naturecnn_backward: bad obs_format %d", obs_format);
    const NatureLayout L(A);
    const NatureActs Q(n, obs_format == B200RL_OBS_U8_NCHW);
    const bf16* P = reinterpret_cast<const bf16*>(packed);
    bf16* act = reinterpret_cast<bf16*>(acts);
    cudaStream_t s = (cudaStream_t)stream;
    // uint8 input: forward left the space-to-depth frames of this minibatch in the workspace
    const bf16* x0 = obs_format == B200RL_OBS_U8_NCHW ? act + Q.x0 : reinterpret_cast<const bf16*>(obs);
    const int64_t* x0rows = obs_format == B200RL_OBS_U8_NCHW ? nullptr : rows;
    // workspace split: [wgrad partials | small partials]
    size_t big = 0;
    {
        auto mx = [&](size_t v) { if (v > big) big = v; };
        mx((size_t)wgrad_plan(n * 512, kC1Ctas, 128).splits * 256 * 64 * 4);
        mx((size_t)wgrad_plan(n * 100, kC2Ctas, 128).splits * 512 * 64 * 4);
        mx((size_t)wgrad_plan(n * 81, kC3Ctas, 128).splits * 640 * 64 * 4);
        mx((size_t)wgrad_plan(n, kFcSplits, 64).splits * 512 * (13 * 256) * 4);
        big = (big + 255) & ~(size_t)255;
    }
    float* wsbig = reinterpret_cast<float*>(workspace);
    float* wssmall = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + big);
    int rc;
    const int A1 = A + 1;
    // ---- heads: dW, db, then dhid_pre = (dhead . Wh) * (hid > 0)
    {
        const int64_t rpb = heads_rows_per_block(n);
        const int nb = (int)ceil_div(n, rpb);
        ProfScope ps(s, "heads_bwd", 4.0 * n * 512 * A1, (double)n * (2048 + 64 + 8 * A1));
        const size_t sd = (size_t)rpb * A1 * sizeof(float);
        if (A1 <= 8) tc_heads_bwd_weight<8><<<nb, 512, sd, s>>>(dhead, act + Q.hid, n, A1, 512, rpb, wssmall);
        else tc_heads_bwd_weight<kMaxHeads><<<nb, 512, sd, s>>>(dhead, act + Q.hid, n, A1, 512, rpb, wssmall);
        tc_heads_fold<<<(unsigned)ceil_div(A1 * 514, 32), 256, 0, s>>>(wssmall, 2 * nb, A1, 512, grads + L.hw, grads + L.hb);
        int db_blocks = (int)ceil_div(n * 64, 256); if (db_blocks > num_sms() * 8) db_blocks = num_sms() * 8;
        tc_heads_bwd_data<<<db_blocks, 256, (size_t)A1 * 2048, s>>>(dhead, params + L.hw, reinterpret_cast<const uint8_t*>(act + Q.m4), n, A1, 512, act + Q.dhid);
        if ((rc = check_launch("naturecnn/heads_bwd", 3))) return rc;
    }
    KGemmParams p;
    // ---- fc: dW[o][c*49+p] = sum_m dhid[m][o] * act3[m][p*64+c]
    {
        const WPlan pl = wgrad_plan(n, kFcSplits, 64);
        { ProfScope ps(s, "fc_wgrad", 2.0 * n * 512 * 3136, (double)n * (3136 + 512) * 2 + 512.0 * 3136 * 4);
          CUtensorMap tmX, tmY;
          if ((rc = make_tmap_2d(&tmX, act + Q.dhid, n, 512, 64, "naturecnn/fc_wgrad"))) return rc;
          if ((rc = make_tmap_2d(&tmY, act + Q.act3, n, 3136, 64, "naturecnn/fc_wgrad"))) return rc;
          const size_t smem = (size_t)4 * (2 + 4) * 64 * 128 + 1024;
          static bool attr_done = false;
          if (!attr_done) {
              cudaError_t e = cudaFuncSetAttribute(tc_wgrad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
              if (e != cudaSuccess) return fail(B200RL_ERR_CUDA, "fc_wgrad: smem attribute: %s", cudaGetErrorString(e));
              attr_done = true;
          }
          tc_wgrad_tma<<<dim3(pl.splits, 4, 13), 160, smem, s>>>(tmX, tmY, n, pl.rows_per_cta, 2, 4, wsbig);
          if ((rc = check_launch("naturecnn/fc_wgrad"))) return rc; }
        { ProfScope ps(s, "wgrad_fold_bias", 0, 0);
          note_launches(1); tc_fold_fc<<<(unsigned)ceil_div((int64_t)512 * 3136, 256), 256, 0, s>>>(wsbig, pl.splits, 512, 13 * 256, 512, 3136, 64, 49, 1.f, grads + L.fcw);
          if ((rc = colsum(act + Q.dhid, n, 512, 512, wssmall, grads + L.fcb, s))) return rc; }
        // dact3_pre = (dhid . Wfc) * (act3 > 0), written on the 9x9 linear grid and the zero-padded 11x11 grid
        gemm_rowmajor(p, act + Q.dhid, n, 8);
        p.Bw = P + L.wfcdg; p.N = 3136; p.out = act + Q.dact3a; p.out2 = act + Q.dact3b; p.dual_dact3 = 1;
        p.ldo = 3136; p.mask_bits = reinterpret_cast<const uint32_t*>(act + Q.m3);
        { ProfScope ps(s, "fc_dgrad", 2.0 * n * 512 * 3136, (double)n * ((3136 + 512) * 2 + 392) + 512.0 * 3136 * 2);
          if ((rc = launch_gemm_tma<256, 4>(p, s, "naturecnn/fc_dgrad"))) return rc; }
    }
    WGradWinParams gw;
    WinParams wp;
    FoldWin fw;
    // ---- conv3: dW from act2 windows x dact3 (9x9 grid), then dact2 = full correlation of padded dact3 with W3
    {
        wgw_defaults(gw);
        gw.X = act + Q.act2; gw.M = n * 81; gw.n = (int)n; gw.G = 81; gw.cpr = 1; gw.nslots = 10; gw.WRX = round8(128 + 20);
        for (int t = 0; t < 9; ++t) gw.shift[t] = (t / 3) * 9 + (t % 3);
        const int st[10] = {0, 1, 2, 3, 4, 5, 6, 7, 7, 8};      // slot 8 duplicates tap 7 so that tap 8 has a partner
        for (int k = 0; k < 10; ++k) { gw.slot_tap[k] = st[k]; gw.slot_cc[k] = 0; }
        gw.Y = act + Q.dact3a; gw.ldy = 64; gw.ncolsY = 64;
        const WPlan pl = wgrad_plan(n * 81, kC3Ctas, 128);
        gw.rows_per_cta = pl.rows_per_cta; gw.ws = wsbig; gw.wsb = wssmall;
        { ProfScope ps(s, "conv3_wgrad", 2.0 * n * 49 * 64 * 576, (double)n * (5184 + 5184) * 2);
          if ((rc = launch_wgrad_win(gw, pl.splits, s, "naturecnn/conv3_wgrad"))) return rc; }
        memset(&fw, 0, sizeof(fw));
        fw.layer = 3; fw.S = pl.splits; fw.nslots = 10; fw.Cout = 64; fw.scale = 1.f;
        for (int k = 0; k < 10; ++k) { fw.slot_tap[k] = st[k]; fw.slot_skip[k] = (k == 8); }
        { ProfScope ps(s, "wgrad_fold_bias", 0, 0);
          fw.wsb = wssmall; fw.db = grads + L.c3b;
          tc_fold_win<<<(unsigned)ceil_div(640 * 64 + 64, 32), 256, 0, s>>>(wsbig, fw, grads + L.c3w);
          if ((rc = check_launch("naturecnn/conv3_fold"))) return rc; }
        win_defaults(wp);
        wp.A = act + Q.dact3b; wp.n = (int)n; wp.G = 121; wp.Wp = 11; wp.M = n * 121; wp.ntaps = 9;
        for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) wp.shift[ky * 3 + kx] = (2 - ky) * 11 + (2 - kx);
        wp.WR = round8(128 + 24);
        wp.Bw = P + L.w3dg; wp.N = 64; wp.vH = 9; wp.vW = 9; wp.out_mode = WOUT_DACT2;
        wp.out = act + Q.dact2a; wp.out2 = act + Q.dact2b; wp.mask_bits = reinterpret_cast<const uint32_t*>(act + Q.m2);
        { ProfScope ps(s, "conv3_dgrad", 2.0 * n * 81 * 64 * 576, (double)n * ((7744 + 6400 + 7744) * 2 + 648));
          if ((rc = launch_conv_win<64, 1, 6, 9>(wp, s, "naturecnn/conv3_dgrad"))) return rc; }
    }
    // ---- conv2: dW from act1 cell windows x dact2 (10x10 grid); dact1 = one N=128 GEMM over the 4 stride-parity
    //      classes (the 4 channel groups of a cell)
    {
        wgw_defaults(gw);
        gw.X = act + Q.act1; gw.M = n * 100; gw.n = (int)n; gw.G = 100; gw.cpr = 2; gw.nslots = 8; gw.WRX = round8(128 + 11);
        gw.shift[0] = 0; gw.shift[1] = 1; gw.shift[2] = 10; gw.shift[3] = 11;
        for (int k = 0; k < 8; ++k) { gw.slot_tap[k] = k >> 1; gw.slot_cc[k] = k & 1; }
        gw.Y = act + Q.dact2a; gw.ldy = 64; gw.ncolsY = 64;
        const WPlan pl = wgrad_plan(n * 100, kC2Ctas, 128);
        gw.rows_per_cta = pl.rows_per_cta; gw.ws = wsbig; gw.wsb = wssmall;
        { ProfScope ps(s, "conv2_wgrad", 2.0 * n * 81 * 64 * 512, (double)n * (12800 + 6400) * 2);
          if ((rc = launch_wgrad_win(gw, pl.splits, s, "naturecnn/conv2_wgrad"))) return rc; }
        memset(&fw, 0, sizeof(fw));
        fw.layer = 2; fw.S = pl.splits; fw.nslots = 8; fw.Cout = 64; fw.scale = 1.f;
        for (int k = 0; k < 8; ++k) { fw.slot_tap[k] = k >> 1; fw.slot_cc[k] = k & 1; }
        { ProfScope ps(s, "wgrad_fold_bias", 0, 0);
          fw.wsb = wssmall; fw.db = grads + L.c2b;
          tc_fold_win<<<(unsigned)ceil_div(512 * 64 + 64, 32), 256, 0, s>>>(wsbig, fw, grads + L.c2w);
          if ((rc = check_launch("naturecnn/conv2_fold"))) return rc; }
        win_defaults(wp);
        wp.A = act + Q.dact2b; wp.n = (int)n; wp.G = 121; wp.Wp = 11; wp.M = n * 121; wp.ntaps = 4;
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) wp.shift[a * 2 + b] = (1 - a) * 11 + (1 - b);
        wp.WR = round8(128 + 12);
        wp.Bw = P + L.w2dg; wp.N = 128; wp.vH = 10; wp.vW = 10; wp.out_mode = WOUT_DACT1;
        wp.out = act + Q.dact1; wp.mask_bits = reinterpret_cast<const uint32_t*>(act + Q.m1);
        { ProfScope ps(s, "conv2_dgrad", 2.0 * n * 400 * 32 * 256, (double)n * ((7744 + 14112) * 2 + 1600));
          if ((rc = launch_conv_win<128, 1, 6, 4>(wp, s, "naturecnn/conv2_dgrad"))) return rc; }
    }
    // ---- conv1 (no data gradient: the input is the observation)
    {
        wgw_defaults(gw);
        gw.X = x0; gw.rows = x0rows; gw.M = n * 512; gw.n = (int)n; gw.G = 441; gw.cpr = 1; gw.nslots = 4; gw.WRX = round8(128 + 22);
        gw.tpi_shift = 2; gw.n_images = x0rows ? (int64_t)1 << 24 : n;       // 4 steps of 128 grid rows per image
        gw.shift[0] = 0; gw.shift[1] = 1; gw.shift[2] = 21; gw.shift[3] = 22;
        for (int k = 0; k < 4; ++k) { gw.slot_tap[k] = k; gw.slot_cc[k] = 0; }
        gw.Y = act + Q.dact1; gw.ldy = 32; gw.ncolsY = 32;
        const WPlan pl = wgrad_plan(n * 512, kC1Ctas, 128);
        gw.rows_per_cta = pl.rows_per_cta; gw.ws = wsbig; gw.wsb = wssmall;
        { ProfScope ps(s, "conv1_wgrad", 2.0 * n * 400 * 32 * 256, (double)n * (28224 + 14112) * 2);
          if ((rc = launch_wgrad_win(gw, pl.splits, s, "naturecnn/conv1_wgrad"))) return rc; }
        memset(&fw, 0, sizeof(fw));
        fw.layer = 1; fw.S = pl.splits; fw.nslots = 4; fw.Cout = 32; fw.scale = 1.0f / 255.0f;
        for (int k = 0; k < 4; ++k) fw.slot_tap[k] = k;
        { ProfScope ps(s, "wgrad_fold_bias", 0, 0);
          fw.wsb = wssmall; fw.db = grads + L.c1b;
          tc_fold_win<<<(unsigned)ceil_div(256 * 32 + 32, 32), 256, 0, s>>>(wsbig, fw, grads + L.c1w);
          if ((rc = check_launch("naturecnn/conv1_fold"))) return rc; }
    }
    return B200RL_OK;
}

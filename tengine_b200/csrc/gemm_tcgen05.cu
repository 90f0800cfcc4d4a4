// gemm_tcgen05.cu -- int8 x int8 -> int32 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// SASS UTCIMMA) with TMA-staged operands, TMEM accumulators and the fused requantising epilogue.
//
//   out[m][oc] = requant( sum_k A[m][k] * B[oc][k] ),   A = NHWC activations (row = pixel, K = padded Cin),
//                                                        B = pre-packed weights [OCp][K], both K-major.
// This is the device's form of the reference's im2col + sgemm_i8 + sgemm_int8 epilogue for 1x1 convolutions
// (source/device/cpu/op/conv/x86/conv_kernel_x86.c:187-242, 1008-1631, 1796-1893) and of ref_fc_int8
// (fc/fc_ref.c:209-297): with NHWC activations a 1x1 convolution IS this GEMM, no im2col pass exists.
//
// Structure (one persistent CTA per SM, 576 threads):
//   warp 0    : TMA producer  (cp.async.bulk.tensor.2d -> 128B/64B/32B-swizzled smem ring, mbarrier expect_tx)
//   warp 1    : MMA issuer    (one elected lane: tcgen05.mma.cta_group::1.kind::i8, 128 x BN x 32 per instruction;
//                              tcgen05.commit releases smem stages / publishes the accumulator)
//   warps 2-17: epilogue      (four warps per TMEM lane quarter, 16-column chunks dealt round-robin: tcgen05.ld 32x32b ->
//                              registers -> requant_fast4 -> int8 -> padded smem tile -> coalesced 16-byte global stores)
// Two TMEM accumulator stages (2 x BN columns) let the MMAs of tile i+1 overlap the epilogue of tile i.
// These layers are HBM-bound (K is 32..1024): the budget is ~8-10 issued instructions per output element, which is
// why the epilogue uses the ~10-instruction requant_fast (common.cuh) and keeps per-channel constants in smem.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

#include <cstdlib>

namespace tb200 {

static constexpr int BLOCK_M = 128;
static constexpr int EPI_WARPS = 16; // four per TMEM lane quarter
static constexpr int EPI_THREADS = EPI_WARPS * 32;
static constexpr int GEMM_THREADS = 64 + EPI_THREADS;
static constexpr int OUT_PAD = 16; // row padding of the staged output tile (bank-conflict-free 16-byte accesses)
static constexpr int MAX_STAGES = 8;

// K-major operand tile in shared memory, rows of `swizzle` bytes, 8-row groups `8*swizzle` bytes apart
// (cute/atom/mma_traits_sm100.hpp: canonical layout ((8,n),2):((swizzle/16,SBO),1), LBO = 1, version 1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, int swizzle)
{
    const uint64_t layout = (swizzle == 128) ? 2ull : (swizzle == 64) ? 4ull : 6ull;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(((8 * swizzle) >> 4) & 0x3fff) << 32; // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}

// UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): S32 accumulate, A/B int8 or uint8,
// both K-major, M = 128, N = block_n.
__host__ __device__ inline uint32_t make_idesc_i8(int block_n, bool a_signed, bool b_signed)
{
    uint32_t d = 0;
    d |= 2u << 4;                        // c_format = S32
    d |= (a_signed ? 1u : 0u) << 7;      // a_format
    d |= (b_signed ? 1u : 0u) << 10;     // b_format
    d |= (uint32_t)(block_n >> 3) << 17; // n_dim
    d |= (uint32_t)(BLOCK_M >> 4) << 24; // m_dim
    return d;
}

struct GemmArgs
{
    long long m;
    int m_tiles, num_super; // 32-bit on purpose: 64-bit divisions in the tile decode cost ~100 instructions each
    int k_blocks, n_tiles, block_n, block_k, stages, swizzle;
    int mt;     // m-tiles (128 rows each) per accumulator stage
    uint32_t nch_rcp, bw_rcp, bh_rcp; // ceil(65536/d): q = (x * rcp) >> 16 is exact for x < 4096, d <= 256
    int oc, ocp, ldo;
    uint32_t idesc;
    uint32_t tmem_cols;
    // conv mode (implicit GEMM): an m-tile is a bw x bh x bn patch of output pixels, a k-block is (tap, channel block)
    int conv, cblocks, kw_n, pad_h, pad_w, cstride, cp;
    int bw, bh, bn, tiles_w, tiles_h, oh, ow, nimg;
    uint32_t a_tx_bytes; // bytes one A load delivers (block_k * rows of the patch)
    // uint8: the B tile carries 16 extra rows, row block_n = all ones, so accumulator column block_n = sum_k x (per pixel)
    int u8, bnx, taps, in_h, in_w;
    int direct_store; // 1: each lane writes its 16 output bytes straight to global memory (no smem staging / copy-out)
    const int32_t* btab; // [taps][OCp]: zx * (sum_c w[oc][tap][c] - Cin*zw), the correction a padding tap needs
};

// m-tile -> first output pixel coordinates (conv mode)
__device__ __forceinline__ void tile_origin(const GemmArgs& g, int mt, int& n0, int& oh0, int& ow0)
{
    const int r = mt / g.tiles_w;
    ow0 = (mt - r * g.tiles_w) * g.bw;
    const int nn = r / g.tiles_h;
    oh0 = (r - nn * g.tiles_h) * g.bh;
    n0 = nn * g.bn;
}

// row r of m-tile mt -> linear output pixel index, or -1 when the row is padding of the tile
__device__ __forceinline__ long long row_pixel(const GemmArgs& g, int mt, int r)
{
    if (!g.conv)
    {
        const long long px = (long long)mt * BLOCK_M + r;
        return px < g.m ? px : -1;
    }
    int n0, oh0, ow0;
    tile_origin(g, mt, n0, oh0, ow0);
    const int t = (int)(((uint32_t)r * g.bw_rcp) >> 16), w = r - t * g.bw;
    const int n = (int)(((uint32_t)t * g.bh_rcp) >> 16), h = t - n * g.bh;
    if (n >= g.bn || n0 + n >= g.nimg || oh0 + h >= g.oh || ow0 + w >= g.ow) return -1;
    return ((long long)(n0 + n) * g.oh + oh0 + h) * g.ow + ow0 + w;
}

struct __align__(16) GemmSmemCtl
{
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
};

__device__ __forceinline__ void quarter_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// 16 accumulator columns of one row -> 16 output bytes staged in shared memory
template <bool FUSE>
__device__ __forceinline__ void epilogue_unit(const uint32_t (&v)[16], uint32_t par_addr, uint32_t dst_addr, uint8_t* gdst, int oc0, int oc_limit,
                                              const EpiParams& e)
{
    uint32_t w[4];
    if (e.fast_ok)
    {
        uint32_t bad = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const float4 p01 = lds_f4(par_addr + j * 32);
            const float4 p23 = lds_f4(par_addr + j * 32 + 16);
            const float m4[4] = {p01.x, p01.z, p23.x, p23.z};
            const int32_t b4[4] = {__float_as_int(p01.y), __float_as_int(p01.w), __float_as_int(p23.y), __float_as_int(p23.w)};
            const int32_t a4[4] = {(int32_t)v[j * 4], (int32_t)v[j * 4 + 1], (int32_t)v[j * 4 + 2], (int32_t)v[j * 4 + 3]};
            w[j] = FUSE ? requant_fast4<false, true>(a4, e, m4, b4, bad, 1u << (4 * j)) : requant_fast4<false, false>(a4, e, m4, b4, bad, 1u << (4 * j));
        }
        if (bad)
        {
            // rare (2.4e-4 of the elements): exact recomputation; fully unrolled so v[] stays in registers
#pragma unroll
            for (int k = 0; k < 16; k++)
                if ((bad >> k) & 1u) w[k >> 2] = requant_fix_byte(w[k >> 2], k & 3, (int32_t)v[k], oc0 + k, e);
        }
    }
    else
    {
        // degenerate scales: literal arithmetic for every element
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            if ((k & 3) == 0) w[k >> 2] = 0;
            if (oc0 + k < oc_limit) w[k >> 2] |= ((uint32_t)requant((int32_t)v[k], oc0 + k, e) & 0xffu) << (8 * (k & 3));
        }
    }
    if (gdst) *reinterpret_cast<uint4*>(gdst) = make_uint4(w[0], w[1], w[2], w[3]);
    else sts_u4(dst_addr, w[0], w[1], w[2], w[3]);
}

// uint8 flavour: v = sum x*w over in-bounds taps (raw bytes), sx = sum x.  The true accumulator is
//   sum (x-zx)(w-zw) = v - zw*sx + corr[oc] + sum_{padding taps t} btab[t][oc]
// with corr[oc] = -zx*sum_k w + taps*Cin*zx*zw (interior pixels) folded into the per-channel constants.
__device__ __forceinline__ void epilogue_unit_u8(const uint32_t (&v)[16], int32_t sx, uint64_t pad_mask, const GemmArgs& g, uint32_t par_addr,
                                                 uint32_t dst_addr, uint8_t* gdst, int oc0, const EpiParams& e)
{
    const int32_t rowc = -e.w_zero * sx;
    int32_t a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = (int32_t)v[k] + rowc;
    if (pad_mask)
    {
        // border pixel: add the per-tap corrections of the taps that fell into the padding (rare rows)
        for (int t = 0; t < g.taps; t++)
            if ((pad_mask >> t) & 1ull)
            {
                const int32_t* bt = g.btab + (size_t)t * g.ocp + oc0;
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if (oc0 + k < g.ocp) a[k] += __ldg(bt + k);
            }
    }
    uint32_t w[4];
    if (e.fast_ok)
    {
        uint32_t bad = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const float4 p01 = lds_f4(par_addr + j * 32);
            const float4 p23 = lds_f4(par_addr + j * 32 + 16);
            const float m4[4] = {p01.x, p01.z, p23.x, p23.z};
            const int32_t z4[4] = {0, 0, 0, 0};
            // corr[oc] travels in the .y lanes
            a[j * 4 + 0] += __float_as_int(p01.y), a[j * 4 + 1] += __float_as_int(p01.w);
            a[j * 4 + 2] += __float_as_int(p23.y), a[j * 4 + 3] += __float_as_int(p23.w);
            const int32_t a4[4] = {a[j * 4], a[j * 4 + 1], a[j * 4 + 2], a[j * 4 + 3]};
            w[j] = requant_fast4<true>(a4, e, m4, z4, bad, 1u << (4 * j));
        }
        // pad lanes of uint8 tensors hold 0, not the zero point
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (oc0 + k >= g.oc) w[k >> 2] &= ~(0xffu << (8 * (k & 3))), bad &= ~(1u << k);
        if (bad)
        {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if ((bad >> k) & 1u) w[k >> 2] = requant_fix_byte(w[k >> 2], k & 3, a[k], oc0 + k, e);
        }
    }
    else
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            if ((k & 3) == 0) w[k >> 2] = 0;
            if (oc0 + k < g.oc)
            {
                const float4 pp = lds_f4(par_addr + (k >> 1) * 16);
                const int32_t corr = __float_as_int((k & 1) ? pp.w : pp.y);
                w[k >> 2] |= ((uint32_t)requant(a[k] + corr, oc0 + k, e) & 0xffu) << (8 * (k & 3));
            }
        }
    }
    if (gdst) *reinterpret_cast<uint4*>(gdst) = make_uint4(w[0], w[1], w[2], w[3]);
    else sts_u4(dst_addr, w[0], w[1], w[2], w[3]);
}

// taps of output pixel (oh, ow) that fall outside the image (bit t = kh*kw_n + kw)
__device__ __forceinline__ uint64_t padding_taps(const GemmArgs& g, int mt, int r)
{
    if (!g.conv || g.taps == 1 && g.pad_h == 0 && g.pad_w == 0) return 0;
    int n0, oh0, ow0;
    tile_origin(g, mt, n0, oh0, ow0);
    const int t = (int)(((uint32_t)r * g.bw_rcp) >> 16), w = r - t * g.bw;
    const int n = (int)(((uint32_t)t * g.bh_rcp) >> 16), h = t - n * g.bh;
    const int iy0 = (oh0 + h) * g.cstride - g.pad_h, ix0 = (ow0 + w) * g.cstride - g.pad_w;
    const int khn = g.taps / g.kw_n;
    if (iy0 >= 0 && ix0 >= 0 && iy0 + khn <= g.in_h && ix0 + g.kw_n <= g.in_w) return 0; // interior
    uint64_t m = 0;
    for (int kh = 0; kh < khn; kh++)
        for (int kw = 0; kw < g.kw_n; kw++)
            if (iy0 + kh < 0 || iy0 + kh >= g.in_h || ix0 + kw < 0 || ix0 + kw >= g.in_w) m |= 1ull << (kh * g.kw_n + kw);
    return m;
}

template <bool U8>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_i8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           uint8_t* __restrict__ out, const GemmArgs g, const __grid_constant__ EpiParams e)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // operand ring first (1024-byte aligned for the 128B swizzle), then the control block, the per-quarter epilogue
    // constants and the per-quarter output staging tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = BLOCK_M * g.block_k, b_bytes = g.bnx * g.block_k;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
    GemmSmemCtl* ctl = reinterpret_cast<GemmSmemCtl*>(smem + (size_t)g.stages * stage_bytes);
    const int opitch = g.block_n + OUT_PAD;
    const uint32_t par_base = smem_u32(ctl) + (uint32_t)sizeof(GemmSmemCtl);   // 4 x [block_n] float2 (m, bias)
    const uint32_t ost_base = par_base + 4u * (uint32_t)g.block_n * 8u;           // 4 x [mt*32][block_n + OUT_PAD]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int acc_cols = g.mt * g.bnx; // TMEM columns of one accumulator stage

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < g.stages; s++) mbar_init(&ctl->full[s], 1), mbar_init(&ctl->empty[s], 1);
        for (int s = 0; s < 2; s++) mbar_init(&ctl->tmem_full[s], 1), mbar_init(&ctl->tmem_empty[s], EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2)
    {
        // TMEM allocation: one warp, power-of-two columns >= 32; base address is written to shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)),
                     "r"(g.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = ctl->tmem_base;

    if (warp == 0)
    {
        // ===================== TMA producer =====================
        if (lane == 0)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (int st = blockIdx.x; st < g.num_super; st += gridDim.x)
            {
                const int msup = st / g.n_tiles;
                const int mt0 = msup * g.mt;
                const int ntile = st - msup * g.n_tiles;
                const int n0 = ntile * g.bnx; // row of this N tile in the packed weight matrix
                for (int i = 0; i < g.mt && mt0 + i < g.m_tiles; i++)
                {
                    const int m0 = (mt0 + i) * BLOCK_M;
                    int cn0 = 0, coh0 = 0, cow0 = 0;
                    if (g.conv) tile_origin(g, mt0 + i, cn0, coh0, cow0);
                    for (int kb = 0; kb < g.k_blocks; kb++)
                    {
                        mbar_wait(&ctl->empty[stage], phase ^ 1);
                        uint8_t* sa = smem + (size_t)stage * stage_bytes;
                        if (!g.conv)
                        {
                            mbar_expect_tx(&ctl->full[stage], a_bytes + b_bytes);
                            tma_load_2d(&tmap_a, &ctl->full[stage], sa, kb * g.block_k, m0);
                            tma_load_2d(&tmap_b, &ctl->full[stage], sa + a_bytes, kb * g.block_k, n0);
                        }
                        else
                        {
                            // k-block = (filter tap, channel block): the A tile is the output patch shifted by the tap;
                            // coordinates outside the image are zero-filled by the TMA unit = the convolution's padding
                            const int tap = kb / g.cblocks, cb = kb - tap * g.cblocks;
                            const int kh = tap / g.kw_n, kw = tap - kh * g.kw_n;
                            mbar_expect_tx(&ctl->full[stage], g.a_tx_bytes + b_bytes);
                            tma_load_4d(&tmap_a, &ctl->full[stage], sa, cb * g.block_k, cow0 * g.cstride - g.pad_w + kw,
                                        coh0 * g.cstride - g.pad_h + kh, cn0);
                            tma_load_2d(&tmap_b, &ctl->full[stage], sa + a_bytes, tap * g.cp + cb * g.block_k, n0);
                        }
                        if (++stage == g.stages) stage = 0, phase ^= 1;
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ===================== MMA issuer =====================
        if (lane == 0)
        {
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            for (int st = blockIdx.x; st < g.num_super; st += gridDim.x)
            {
                const int mt0 = (st / g.n_tiles) * g.mt;
                mbar_wait(&ctl->tmem_empty[as], aphase ^ 1); // the epilogue has drained this accumulator stage
                tcgen05_fence_after();
                for (int i = 0; i < g.mt && mt0 + i < g.m_tiles; i++)
                {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * acc_cols + i * g.bnx);
                    for (int kb = 0; kb < g.k_blocks; kb++)
                    {
                        mbar_wait(&ctl->full[stage], phase); // TMA bytes have landed
                        tcgen05_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                        const uint64_t da = make_smem_desc(sa, g.swizzle), db = make_smem_desc(sa + a_bytes, g.swizzle);
                        for (int k = 0; k < g.block_k / 32; k++)
                        {
                            // advance 32 bytes (one UMMA_K of int8) inside the swizzled row: +2 in 16-byte units
                            umma_i8(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), g.idesc, (kb | k) ? 1u : 0u);
                        }
                        tcgen05_commit(&ctl->empty[stage]); // smem stage reusable once these MMAs have read it
                        if (++stage == g.stages) stage = 0, phase ^= 1;
                    }
                }
                tcgen05_commit(&ctl->tmem_full[as]); // all m-tiles of this accumulator stage are complete
                if (++as == 2) as = 0, aphase ^= 1;
            }
        }
    }
    else
    {
        // ===================== epilogue (warps 2..17) =====================
        // TMEM lane quarter q (hardware rule: a warp may only touch lanes 32*(warp_id % 4)...) is served by the four
        // warps {q, q+4, q+8, q+12}: they deal the (m-tile, 16-column chunk) units of the stage round-robin, stage the
        // requantised bytes of "their" rows in smem and copy them out with coalesced 16-byte stores.  Only these 128
        // threads synchronise with each other (named barriers); nothing here is block-wide.
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int tq = sub * 32 + lane; // 0..127 inside the quarter group
        const int nch = g.block_n >> 4; // 16-column chunks per m-tile
        const uint32_t par_s = par_base + (uint32_t)(q * g.block_n * 8);
        const uint32_t ost_s = ost_base + (uint32_t)(q * g.mt * 32 * opitch);
        int as = 0;
        uint32_t aphase = 0;
        int loaded_n0 = -1;
        for (int st = blockIdx.x; st < g.num_super; st += gridDim.x)
        {
            const int msup = st / g.n_tiles;
            const int mt0 = msup * g.mt;
            const int n0 = (st - msup * g.n_tiles) * g.block_n;
            const int rem = g.m_tiles - mt0;
            const int mtc = rem < g.mt ? rem : g.mt;
            if (n0 != loaded_n0)
            {
                // per-channel fast-path constants (m, bias) of this N tile; pad / overhanging channels get (0, 0)
                for (int c = tq; c < g.block_n; c += 128)
                {
                    const int oc = n0 + c;
                    sts_f2(par_s + c * 8, (oc < g.ocp && e.fast_ok) ? __ldg(e.fast_par + oc) : make_float2(0.f, 0.f));
                }
                loaded_n0 = n0;
            }
            mbar_wait(&ctl->tmem_full[as], aphase);
            tcgen05_fence_after();
            quarter_bar_sync(1 + q); // A: constants visible; the group finished copying the previous stage out
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * acc_cols);
            // this warp's units: u = sub, sub+4, ... ; unit u = (m-tile i = u / nch, chunk ci = u % nch), tracked incrementally
            uint32_t va[16], vb[16];
            int i = 0, ci = sub;
            while (ci >= nch) ci -= nch, i++;
            auto process = [&](const uint32_t (&v)[16], int ii, int cc)
            {
                const int c = cc * 16;
                uint8_t* gdst = nullptr;
                bool skip = false;
                if (g.direct_store)
                {
                    const long long px = row_pixel(g, mt0 + ii, q * 32 + lane);
                    skip = px < 0 || n0 + c >= g.ocp;
                    gdst = out + (size_t)(px < 0 ? 0 : px) * g.ldo + n0 + c;
                }
                int32_t sx = 0;
                if (U8) // warp-collective TMEM load: before any lane-dependent branch
                {
                    sx = (int32_t)tmem_ld1(tbase + ii * g.bnx + g.block_n);
                    tmem_ld_wait();
                }
                const uint32_t sdst = ost_s + (uint32_t)((ii * 32 + lane) * opitch + c);
                if (skip)
                    ;
                else if (U8)
                    epilogue_unit_u8(v, sx, padding_taps(g, mt0 + ii, q * 32 + lane), g, par_s + c * 8, sdst, gdst, n0 + c, e);
                else if (e.fuse_bias)
                    epilogue_unit<true>(v, par_s + c * 8, sdst, gdst, n0 + c, g.oc, e);
                else
                    epilogue_unit<false>(v, par_s + c * 8, sdst, gdst, n0 + c, g.oc, e);
            };
            if (i < mtc) tmem_ld16(tbase + i * g.bnx + ci * 16, va);
            while (i < mtc)
            {
                tmem_ld_wait();
                int i2 = i, c2 = ci + 4;
                while (c2 >= nch) c2 -= nch, i2++;
                if (i2 < mtc) tmem_ld16(tbase + i2 * g.bnx + c2 * 16, vb); // next unit's accumulators are in flight ...
                process(va, i, ci);                                          // ... while this unit is requantised
                i = i2, ci = c2;
                if (i >= mtc) break;
                tmem_ld_wait();
                i2 = i, c2 = ci + 4;
                while (c2 >= nch) c2 -= nch, i2++;
                if (i2 < mtc) tmem_ld16(tbase + i2 * g.bnx + c2 * 16, va);
                process(vb, i, ci);
                i = i2, ci = c2;
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->tmem_empty[as]); // accumulator drained: the MMA warp may overwrite it
            if (++as == 2) as = 0, aphase ^= 1;
            if (g.direct_store) continue; // bytes already went to global memory
            quarter_bar_sync(5 + q); // B: the quarter's staged rows are complete
            // coalesced copy-out of this quarter's rows: consecutive threads write consecutive 16-byte pieces
            const int total_vec = mtc * 32 * nch;
            for (int vi = tq; vi < total_vec; vi += 128)
            {
                const int lr = (int)(((uint32_t)vi * g.nch_rcp) >> 16);
                const int cv = vi - lr * nch;
                const long long grow = row_pixel(g, mt0 + (lr >> 5), q * 32 + (lr & 31));
                if (grow >= 0 && n0 + cv * 16 < g.ocp)
                    *reinterpret_cast<uint4*>(out + (size_t)grow * g.ldo + n0 + cv * 16) = lds_u4(ost_s + (uint32_t)(lr * opitch + cv * 16));
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2)
    {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(g.tmem_cols) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

int tmap_encode(void* tmap, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                const uint32_t* elem_strides, int swizzle_bytes)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return TB200_ERR_CUDA;
    cuuint64_t d[5], st[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) d[i] = dims[i], bx[i] = box[i], es[i] = elem_strides ? elem_strides[i] : 1;
    for (int i = 0; i + 1 < rank; i++) st[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                            : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                            : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                  : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc((CUtensorMap*)tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : TB200_ERR_CUDA;
}

static int encode_2d(void* tmap, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_inner,
                     uint32_t box_rows, int swizzle)
{
    const uint64_t dims[2] = {inner, rows}, strides[1] = {pitch};
    const uint32_t box[2] = {box_inner, box_rows};
    return tmap_encode(tmap, base, 2, dims, strides, box, nullptr, swizzle);
}

int gemm_block_n(int ocp, int u8)
{
    if (!u8) return ocp <= 256 ? ocp : 128;
    return ocp <= 240 ? ocp : 112; // +16 rows for the ones-row keeps the UMMA N at <= 256 / 128
}

int gemm_plan_create(GemmPlan* p, const void* a, long long lda, const void* b, long long m, int k, int oc, int ocp, int ldo,
                     int variant, int u8)
{
    if (m <= 0 || k <= 0 || (k & 15) || (ocp & 15) || (lda & 15) || (ldo & 15)) return TB200_ERR_INVALID;
    memset(p, 0, sizeof *p);
    p->m = m, p->k = k, p->oc = oc, p->ocp = ocp, p->ldo = ldo, p->variant = variant;
    p->block_k = k <= 32 ? 32 : (k <= 64 ? 64 : 128);
    p->swizzle = p->block_k;
    p->k_blocks = (k + p->block_k - 1) / p->block_k;
    p->u8 = u8;
    p->block_n = gemm_block_n(ocp, u8);
    p->bnx = p->block_n + (u8 ? 16 : 0);
    p->taps = 1;
    p->n_tiles = (ocp + p->block_n - 1) / p->block_n;
    p->m_tiles = (m + BLOCK_M - 1) / BLOCK_M;
    const int a_bytes = BLOCK_M * p->block_k, b_bytes = (p->bnx * p->block_k + 1023) & ~1023;
    // m-tiles per accumulator stage: amortise the per-stage synchronisation over ~256 TMEM columns of work
    p->mt = 1;
    if (p->n_tiles == 1)
        while (p->mt < 4 && 2 * (p->mt * 2) * p->bnx <= 512 && (long long)(p->mt * 2) <= p->m_tiles) p->mt *= 2;
    const int epi_bytes = 4 * p->block_n * 8 + BLOCK_M * p->mt * (p->block_n + OUT_PAD) + 2048;
    int stages = (224 * 1024 - epi_bytes) / (a_bytes + b_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return TB200_ERR_INVALID;
    p->stages = stages;
    int rc = encode_2d(p->tmap_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, p->block_k, BLOCK_M, p->swizzle);
    if (rc) return rc;
    rc = encode_2d(p->tmap_b, b, (uint64_t)k, (uint64_t)p->n_tiles * p->bnx, (uint64_t)k, p->block_k, p->bnx, p->swizzle);
    return rc;
}

// Implicit-GEMM plan for a dense (group 1, dilation 1) convolution with any kernel size and stride 1 or 2:
// A = 4-D tensor map (C, W, H, N) over the NHWC input with traversal strides (1, s, s, 1); B = [OCp][taps*Cp].
int gemm_plan_create_conv(GemmPlan* p, const void* in, const void* w, const ConvShape& s, int u8)
{
    if (s.group != 1 || s.dh != 1 || s.dw != 1 || s.sh != s.sw || (s.sh != 1 && s.sh != 2)) return TB200_ERR_UNSUPPORTED;
    const int taps = s.kh * s.kw;
    if (taps > 1 && (s.cp % 32)) return TB200_ERR_UNSUPPORTED; // a k-block must not straddle two taps
    memset(p, 0, sizeof *p);
    p->conv = 1;
    p->m = (long long)s.n * s.oh * s.ow, p->oc = s.oc, p->ocp = s.ocp, p->ldo = s.ocp, p->variant = 0;
    if (taps == 1) p->block_k = s.cp <= 32 ? 32 : (s.cp <= 64 ? 64 : 128);
    else p->block_k = (s.cp % 128 == 0) ? 128 : ((s.cp % 64 == 0) ? 64 : 32);
    p->swizzle = p->block_k;
    p->cblocks = (s.cp + p->block_k - 1) / p->block_k;
    p->k_blocks = taps * p->cblocks;
    p->k = taps * s.cp;
    p->u8 = u8;
    p->block_n = gemm_block_n(s.ocp, u8);
    p->bnx = p->block_n + (u8 ? 16 : 0);
    p->taps = taps, p->in_h = s.h, p->in_w = s.w;
    if (taps > 64) return TB200_ERR_UNSUPPORTED;
    p->n_tiles = (s.ocp + p->block_n - 1) / p->block_n;
    // output patch of one m-tile: whole rows when they fit (then the patch is contiguous in the NHWC output)
    if (s.ow <= BLOCK_M)
    {
        p->bw = s.ow;
        p->bh = BLOCK_M / s.ow < s.oh ? BLOCK_M / s.ow : s.oh;
        p->bn = (p->bh == s.oh) ? BLOCK_M / (s.ow * s.oh) : 1;
        if (p->bn > s.n) p->bn = s.n;
        if (p->bn < 1) p->bn = 1;
    }
    else
        p->bw = BLOCK_M, p->bh = 1, p->bn = 1;
    p->tiles_w = (s.ow + p->bw - 1) / p->bw;
    p->tiles_h = (s.oh + p->bh - 1) / p->bh;
    const long long tiles_n = (s.n + p->bn - 1) / p->bn;
    p->m_tiles = (long long)p->tiles_w * p->tiles_h * tiles_n;
    p->kw_n = s.kw, p->pad_h = s.ph0, p->pad_w = s.pw0, p->cstride = s.sh, p->cp = s.cp, p->oh = s.oh, p->ow = s.ow, p->nimg = s.n;
    p->a_tx_bytes = (uint32_t)(p->block_k * p->bw * p->bh * p->bn);
    const int a_bytes = BLOCK_M * p->block_k, b_bytes = (p->bnx * p->block_k + 1023) & ~1023;
    p->mt = 1;
    if (p->n_tiles == 1)
        while (p->mt < 4 && 2 * (p->mt * 2) * p->bnx <= 512 && (long long)(p->mt * 2) <= p->m_tiles) p->mt *= 2;
    const int epi_bytes = 4 * p->block_n * 8 + BLOCK_M * p->mt * (p->block_n + OUT_PAD) + 2048;
    int stages = (224 * 1024 - epi_bytes) / (a_bytes + b_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return TB200_ERR_INVALID;
    p->stages = stages;
    const uint64_t dims[4] = {(uint64_t)s.cp, (uint64_t)s.w, (uint64_t)s.h, (uint64_t)s.n};
    const uint64_t strides[3] = {(uint64_t)s.cp, (uint64_t)s.w * s.cp, (uint64_t)s.h * s.w * s.cp};
    const uint32_t box[4] = {(uint32_t)p->block_k, (uint32_t)((p->bw - 1) * s.sw + 1), (uint32_t)((p->bh - 1) * s.sh + 1), (uint32_t)p->bn};
    const uint32_t estr[4] = {1u, (uint32_t)s.sw, (uint32_t)s.sh, 1u};
    if (box[1] > 256 || box[2] > 256 || box[3] > 256) return TB200_ERR_UNSUPPORTED;
    int rc = tmap_encode(p->tmap_a, in, 4, dims, strides, box, estr, p->swizzle);
    if (rc) return rc;
    return encode_2d(p->tmap_b, w, (uint64_t)p->k, (uint64_t)p->n_tiles * p->bnx, (uint64_t)p->k, p->block_k, p->bnx, p->swizzle);
}

cudaError_t launch_gemm_i8(const GemmPlan& p, void* out, const EpiParams& e, const int32_t* btab, int num_sms, cudaStream_t st)
{
    if (p.block_n <= 0 || p.mt <= 0 || p.stages <= 0) return cudaErrorInvalidValue; // plan was never created
    GemmArgs g;
    g.m = p.m, g.m_tiles = (int)p.m_tiles, g.k_blocks = p.k_blocks, g.n_tiles = p.n_tiles, g.block_n = p.block_n;
    g.conv = p.conv, g.cblocks = p.cblocks, g.kw_n = p.kw_n, g.pad_h = p.pad_h, g.pad_w = p.pad_w, g.cstride = p.cstride, g.cp = p.cp;
    g.bw = p.bw, g.bh = p.bh, g.bn = p.bn, g.tiles_w = p.tiles_w, g.tiles_h = p.tiles_h, g.oh = p.oh, g.ow = p.ow, g.nimg = p.nimg;
    g.a_tx_bytes = p.a_tx_bytes;
    g.u8 = p.u8, g.bnx = p.bnx, g.taps = p.taps, g.in_h = p.in_h, g.in_w = p.in_w, g.btab = btab;
    static const int direct_env = getenv("TB200_GEMM_DIRECT_STORE") ? atoi(getenv("TB200_GEMM_DIRECT_STORE")) : 0;
    g.direct_store = direct_env;
    g.mt = p.mt;
    g.num_super = (int)(((p.m_tiles + p.mt - 1) / p.mt) * p.n_tiles);
    g.nch_rcp = (65536u + (uint32_t)(p.block_n >> 4) - 1) / (uint32_t)(p.block_n >> 4);
    g.bw_rcp = p.conv ? (65536u + (uint32_t)p.bw - 1) / (uint32_t)p.bw : 0;
    g.bh_rcp = p.conv ? (65536u + (uint32_t)p.bh - 1) / (uint32_t)p.bh : 0;
    g.block_k = p.block_k, g.stages = p.stages, g.swizzle = p.swizzle, g.oc = p.oc, g.ocp = p.ocp, g.ldo = p.ldo;
    g.idesc = make_idesc_i8(p.bnx, !p.u8, !p.u8);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.mt * p.bnx)) cols <<= 1;
    g.tmem_cols = cols;
    const int a_bytes = BLOCK_M * p.block_k, b_bytes = (p.bnx * p.block_k + 1023) & ~1023;
    const size_t smem = (size_t)p.stages * (a_bytes + b_bytes) + sizeof(GemmSmemCtl) + 4 * (size_t)p.block_n * 8 +
                        (size_t)BLOCK_M * p.mt * (p.block_n + OUT_PAD) + 1024;
    static bool attr_set = false;
    if (!attr_set)
    {
        cudaError_t err = cudaFuncSetAttribute(gemm_i8_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(gemm_i8_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err != cudaSuccess) return err;
        attr_set = true;
    }
    const int grid = (int)(g.num_super < num_sms ? g.num_super : num_sms);
    CUtensorMap ta, tb;
    memcpy(&ta, p.tmap_a, sizeof ta);
    memcpy(&tb, p.tmap_b, sizeof tb);
    if (p.u8) gemm_i8_tcgen05_kernel<true><<<grid, GEMM_THREADS, smem, st>>>(ta, tb, (uint8_t*)out, g, e);
    else gemm_i8_tcgen05_kernel<false><<<grid, GEMM_THREADS, smem, st>>>(ta, tb, (uint8_t*)out, g, e);
    return cudaGetLastError();
}

} // namespace tb200

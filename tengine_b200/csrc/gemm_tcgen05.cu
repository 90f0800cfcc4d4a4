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

namespace tb200 {

static constexpr int BLOCK_M = 128;
static constexpr int EPI_WARPS = 16; // four per TMEM lane quarter
static constexpr int EPI_THREADS = EPI_WARPS * 32;
static constexpr int GEMM_THREADS = 64 + EPI_THREADS;
static constexpr int OUT_PAD = 16; // row padding of the staged output tile (bank-conflict-free 16-byte accesses)
static constexpr int MAX_STAGES = 8;

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (-> CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
    {
        if (clock64() - t0 > 4000000000LL) __trap(); // ~2 s
    }
}
__device__ __forceinline__ void tma_load_2d(const void* tmap, uint64_t* bar, void* smem, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
    long long m, m_tiles;
    int k_blocks, n_tiles, block_n, block_k, stages, swizzle;
    int oc, ocp, ldo;
    uint32_t idesc;
    uint32_t tmem_cols;
};

struct __align__(16) GemmSmemCtl
{
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void epi_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_i8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           uint8_t* __restrict__ out, const GemmArgs g, const EpiParams e)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // operand ring first (1024-byte aligned for the 128B swizzle), control block after it
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = BLOCK_M * g.block_k, b_bytes = g.block_n * g.block_k;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
    GemmSmemCtl* ctl = reinterpret_cast<GemmSmemCtl*>(smem + (size_t)g.stages * stage_bytes);
    float2* epi_par = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(ctl) + sizeof(GemmSmemCtl)); // [block_n] (m, bias)
    uint8_t* ostage = reinterpret_cast<uint8_t*>(epi_par) + (size_t)g.block_n * sizeof(float2);         // [128][block_n + OUT_PAD]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long num_tiles = g.m_tiles * g.n_tiles;

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
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                const int m0 = (int)((tile / g.n_tiles) * BLOCK_M), n0 = (int)(tile % g.n_tiles) * g.block_n;
                for (int kb = 0; kb < g.k_blocks; kb++)
                {
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    mbar_expect_tx(&ctl->full[stage], a_bytes + b_bytes);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    tma_load_2d(&tmap_a, &ctl->full[stage], sa, kb * g.block_k, m0);
                    tma_load_2d(&tmap_b, &ctl->full[stage], sa + a_bytes, kb * g.block_k, n0);
                    if (++stage == g.stages) stage = 0, phase ^= 1;
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
            for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                mbar_wait(&ctl->tmem_empty[as], aphase ^ 1); // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * g.block_n);
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
                tcgen05_commit(&ctl->tmem_full[as]); // accumulator complete
                if (++as == 2) as = 0, aphase ^= 1;
            }
        }
    }
    else
    {
        // ===================== epilogue (warps 2..17) =====================
        const int q = warp & 3;            // TMEM lane quarter this warp may access (hardware rule: warp_id % 4)
        const int split = (warp - 2) >> 2; // 0..3: which of the quarter's four warps
        const int et = threadIdx.x - 64;   // 0..EPI_THREADS-1
        const int opitch = g.block_n + OUT_PAD;
        const int vec_per_row = g.block_n >> 4;
        const uint32_t par_s = smem_u32(epi_par), ost_s = smem_u32(ostage);
        int as = 0;
        uint32_t aphase = 0;
        int loaded_n0 = -1;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        {
            const long long m0 = (tile / g.n_tiles) * BLOCK_M;
            const int n0 = (int)(tile % g.n_tiles) * g.block_n;
            if (n0 != loaded_n0)
            {
                // per-channel fast-path constants (m, bias) of this N tile -> smem; pad / overhanging channels get (0, 0)
                for (int c = et; c < g.block_n; c += EPI_THREADS)
                {
                    const int oc = n0 + c;
                    sts_f2(par_s + c * 8, (oc < g.ocp && e.fast_ok) ? __ldg(e.fast_par + oc) : make_float2(0.f, 0.f));
                }
                loaded_n0 = n0;
            }
            mbar_wait(&ctl->tmem_full[as], aphase);
            tcgen05_fence_after();
            epi_bar_sync(1); // A: constants visible; everybody finished copying the previous tile out of `ostage`
            const int rloc = q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * g.block_n);
            const uint32_t srow = ost_s + (uint32_t)(rloc * opitch);
            for (int c = split * 16; c < g.block_n; c += 16 * (EPI_WARPS / 4))
            {
                uint32_t v[16];
                tmem_ld16(taddr + c, v);
                tmem_ld_wait();
                uint32_t w[4];
                if (e.fast_ok)
                {
                    uint32_t bad = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        const float4 p01 = lds_f4(par_s + (c + j * 4) * 8);
                        const float4 p23 = lds_f4(par_s + (c + j * 4 + 2) * 8);
                        const float m4[4] = {p01.x, p01.z, p23.x, p23.z};
                        const int32_t b4[4] = {__float_as_int(p01.y), __float_as_int(p01.w), __float_as_int(p23.y), __float_as_int(p23.w)};
                        const int32_t a4[4] = {(int32_t)v[j * 4], (int32_t)v[j * 4 + 1], (int32_t)v[j * 4 + 2], (int32_t)v[j * 4 + 3]};
                        w[j] = requant_fast4<false>(a4, e, m4, b4, bad, 1u << (4 * j));
                    }
                    if (bad)
                    {
                        // rare (2.4e-4 of the elements): exact recomputation; fully unrolled so v[] stays in registers
#pragma unroll
                        for (int k = 0; k < 16; k++)
                            if ((bad >> k) & 1u) w[k >> 2] = requant_fix_byte(w[k >> 2], k & 3, (int32_t)v[k], n0 + c + k, e);
                    }
                }
                else
                {
                    // degenerate scales: literal arithmetic for every element
#pragma unroll
                    for (int k = 0; k < 16; k++)
                    {
                        if ((k & 3) == 0) w[k >> 2] = 0;
                        if (n0 + c + k < g.oc) w[k >> 2] |= ((uint32_t)requant((int32_t)v[k], n0 + c + k, e) & 0xffu) << (8 * (k & 3));
                    }
                }
                sts_u4(srow + c, w[0], w[1], w[2], w[3]);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->tmem_empty[as]); // accumulator drained: the MMA warp may overwrite it
            if (++as == 2) as = 0, aphase ^= 1;
            epi_bar_sync(2); // B: the staged tile is complete
            // coalesced copy-out: consecutive threads write consecutive 16-byte pieces of a row, then the next row
            const int total_vec = BLOCK_M * vec_per_row;
            for (int vi = et; vi < total_vec; vi += EPI_THREADS)
            {
                const int r = vi / vec_per_row, cv = vi - r * vec_per_row;
                if (m0 + r < g.m && n0 + cv * 16 < g.ocp)
                    *reinterpret_cast<uint4*>(out + (size_t)(m0 + r) * g.ldo + n0 + cv * 16) = lds_u4(ost_s + (uint32_t)(r * opitch + cv * 16));
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

static int encode_2d(void* tmap, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_inner,
                     uint32_t box_rows, int swizzle)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return TB200_ERR_CUDA;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {pitch};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                           : (swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc((CUtensorMap*)tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : TB200_ERR_CUDA;
}

int gemm_plan_create(GemmPlan* p, const void* a, long long lda, const void* b, long long m, int k, int oc, int ocp, int ldo,
                     int variant)
{
    if (m <= 0 || k <= 0 || (k & 15) || (ocp & 15) || (lda & 15) || (ldo & 15)) return TB200_ERR_INVALID;
    p->m = m, p->k = k, p->oc = oc, p->ocp = ocp, p->ldo = ldo, p->variant = variant;
    p->block_k = k <= 32 ? 32 : (k <= 64 ? 64 : 128);
    p->swizzle = p->block_k;
    p->k_blocks = (k + p->block_k - 1) / p->block_k;
    p->block_n = ocp <= 256 ? ocp : 128;
    if (variant & 0x100) p->block_n = ocp <= 256 ? ocp : 256; // wide-N variant
    p->n_tiles = (ocp + p->block_n - 1) / p->block_n;
    p->m_tiles = (m + BLOCK_M - 1) / BLOCK_M;
    const int a_bytes = BLOCK_M * p->block_k, b_bytes = (p->block_n * p->block_k + 1023) & ~1023;
    const int epi_bytes = p->block_n * 8 + BLOCK_M * (p->block_n + OUT_PAD) + 2048;
    int stages = (224 * 1024 - epi_bytes) / (a_bytes + b_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return TB200_ERR_INVALID;
    p->stages = stages;
    int rc = encode_2d(p->tmap_a, a, (uint64_t)k, (uint64_t)m, (uint64_t)lda, p->block_k, BLOCK_M, p->swizzle);
    if (rc) return rc;
    rc = encode_2d(p->tmap_b, b, (uint64_t)k, (uint64_t)ocp, (uint64_t)k, p->block_k, p->block_n, p->swizzle);
    return rc;
}

cudaError_t launch_gemm_i8(const GemmPlan& p, void* out, const EpiParams& e, int num_sms, cudaStream_t st)
{
    GemmArgs g;
    g.m = p.m, g.m_tiles = p.m_tiles, g.k_blocks = p.k_blocks, g.n_tiles = p.n_tiles, g.block_n = p.block_n;
    g.block_k = p.block_k, g.stages = p.stages, g.swizzle = p.swizzle, g.oc = p.oc, g.ocp = p.ocp, g.ldo = p.ldo;
    g.idesc = make_idesc_i8(p.block_n, !e.is_uint8, !e.is_uint8);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.block_n)) cols <<= 1;
    g.tmem_cols = cols;
    const int a_bytes = BLOCK_M * p.block_k, b_bytes = (p.block_n * p.block_k + 1023) & ~1023;
    const size_t smem = (size_t)p.stages * (a_bytes + b_bytes) + sizeof(GemmSmemCtl) + 16 + (size_t)p.block_n * 8 +
                        (size_t)BLOCK_M * (p.block_n + OUT_PAD) + 1024;
    static bool attr_set = false;
    if (!attr_set)
    {
        cudaError_t err = cudaFuncSetAttribute(gemm_i8_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (err != cudaSuccess) return err;
        attr_set = true;
    }
    long long tiles = p.m_tiles * p.n_tiles;
    const int grid = (int)(tiles < num_sms ? tiles : num_sms);
    CUtensorMap ta, tb;
    memcpy(&ta, p.tmap_a, sizeof ta);
    memcpy(&tb, p.tmap_b, sizeof tb);
    gemm_i8_tcgen05_kernel<<<grid, GEMM_THREADS, smem, st>>>(ta, tb, (uint8_t*)out, g, e);
    return cudaGetLastError();
}

} // namespace tb200

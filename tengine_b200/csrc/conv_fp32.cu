// conv_fp32.cu -- the fp32 members of the path: Winograd F(4x4, 3x3) convolution and depthwise 3x3,
// on NCHW fp32 tensors as Tengine holds them.  The north star's fp32 clause is a parity clause (within 1e-4 relative of the
// reference CPU kernels); no BASELINE.json configuration runs fp32, so these are plain CUDA-core FFMA kernels (a tf32 tensor-core
// contraction would break the 1e-4 bound), reachable through the kernel-level ABI (tb200k_conv_winograd43_f32, tb200k_conv_dw3x3_f32).
//
//   winograd: wino_conv_kernel_x86.c -- conv3x3s1_winograd43_transform_kernel_sse (:1118, G = ktm 6x3 :1123-1124),
//             conv3x3s1_winograd43_sse (:126: input transform B^T d B per 6x6 tile, 36 [OC x C] x [C x tiles] products, output
//             transform A^T m A + bias), wino_conv_hcl_run (:1376; activation afterwards, :1435-1438); eligibility
//             conv_kernel_x86.c:1896-1915 (3x3, stride 1, dilation 1, group 1).
//   depthwise: conv_dw_kernel_x86.c:2524 conv_dw_run (3x3, stride 1 / 2, bias, relu / relu6 fused, :121-2457).
#include <cuda_runtime.h>

#include "kernels.h"

namespace tb200 {

__device__ __forceinline__ float act_f32(float v, int activation)
{
    // relu() of wino_conv_kernel_x86.c / the fused epilogues of conv_dw_kernel_x86.c: 0 -> max(v, 0); > 0 -> clip to [0, activation]
    if (activation >= 0) v = fmaxf(v, 0.f);
    if (activation > 0) v = fminf(v, (float)activation);
    return v;
}

// ---- kernel transform: U[36][OC][C] = G g G^T --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wino43_weight_kernel(const float* __restrict__ w, float* __restrict__ U, int oc, int c)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= oc * c) return;
    const float G[6][3] = {{1.0f / 4, 0.0f, 0.0f}, {-1.0f / 6, -1.0f / 6, -1.0f / 6}, {-1.0f / 6, 1.0f / 6, -1.0f / 6},
                           {1.0f / 24, 1.0f / 12, 1.0f / 6}, {1.0f / 24, -1.0f / 12, 1.0f / 6}, {0.0f, 0.0f, 1.0f}};
    const float* g = w + (size_t)idx * 9;
    float tmp[6][3];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) tmp[i][j] = g[0 * 3 + j] * G[i][0] + g[1 * 3 + j] * G[i][1] + g[2 * 3 + j] * G[i][2];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++)
            U[(size_t)(i * 6 + j) * oc * c + idx] = tmp[i][0] * G[j][0] + tmp[i][1] * G[j][1] + tmp[i][2] * G[j][2];
}

// ---- input transform: V[36][C][T] = B^T d B, one thread per (channel, tile) ----------------------------------------------------
__global__ void __launch_bounds__(256) wino43_input_kernel(const float* __restrict__ in, float* __restrict__ V, int n, int c, int h, int w, int ph, int pw,
                                                           int tiles_h, int tiles_w)
{
    const long long T = (long long)n * tiles_h * tiles_w;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * c) return;
    const int ch = (int)(idx / T);
    const long long t = idx - (long long)ch * T;
    const int tw = (int)(t % tiles_w), th = (int)((t / tiles_w) % tiles_h), img = (int)(t / ((long long)tiles_w * tiles_h));
    const float* src = in + ((size_t)img * c + ch) * h * w;
    float d[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++)
        {
            const int y = th * 4 - ph + i, x = tw * 4 - pw + j;
            d[i][j] = (y >= 0 && y < h && x >= 0 && x < w) ? __ldg(src + (size_t)y * w + x) : 0.f;
        }
    // B^T (itm): rows {4,0,-5,0,1,0} {0,-4,-4,1,1,0} {0,4,-4,-1,1,0} {0,-2,-1,2,1,0} {0,2,-1,-2,1,0} {0,4,0,-5,0,1}
    float m[6][6];
#pragma unroll
    for (int j = 0; j < 6; j++)
    {
        const float r0 = d[0][j], r1 = d[1][j], r2 = d[2][j], r3 = d[3][j], r4 = d[4][j], r5 = d[5][j];
        m[0][j] = 4.f * r0 - 5.f * r2 + r4;
        m[1][j] = -4.f * (r1 + r2) + r3 + r4;
        m[2][j] = 4.f * (r1 - r2) - r3 + r4;
        m[3][j] = -2.f * r1 - r2 + 2.f * r3 + r4;
        m[4][j] = 2.f * r1 - r2 - 2.f * r3 + r4;
        m[5][j] = 4.f * r1 - 5.f * r3 + r5;
    }
#pragma unroll
    for (int i = 0; i < 6; i++)
    {
        const float r0 = m[i][0], r1 = m[i][1], r2 = m[i][2], r3 = m[i][3], r4 = m[i][4], r5 = m[i][5];
        float o[6];
        o[0] = 4.f * r0 - 5.f * r2 + r4;
        o[1] = -4.f * (r1 + r2) + r3 + r4;
        o[2] = 4.f * (r1 - r2) - r3 + r4;
        o[3] = -2.f * r1 - r2 + 2.f * r3 + r4;
        o[4] = 2.f * r1 - r2 - 2.f * r3 + r4;
        o[5] = 4.f * r1 - 5.f * r3 + r5;
#pragma unroll
        for (int j = 0; j < 6; j++) V[((size_t)(i * 6 + j) * c + ch) * T + t] = o[j];
    }
}

// ---- 36 products M[k] = U[k] (OC x C) . V[k] (C x T): 64 x 64 tiles, 16-deep k panels, 4 x 4 outputs per thread ------------------
__global__ void __launch_bounds__(256) wino43_gemm_kernel(const float* __restrict__ U, const float* __restrict__ V, float* __restrict__ M, int oc, int c, long long T)
{
    __shared__ float sU[16][64 + 1];
    __shared__ float sV[16][64];
    const int k = blockIdx.z;
    const float* Uk = U + (size_t)k * oc * c;
    const float* Vk = V + (size_t)k * c * T;
    float* Mk = M + (size_t)k * oc * T;
    const int oc0 = blockIdx.y * 64;
    const long long t0 = (long long)blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4; // 16 x 16 threads, each 4 (oc) x 4 (t)
    float acc[4][4] = {};
    for (int c0 = 0; c0 < c; c0 += 16)
    {
        for (int i = threadIdx.x; i < 16 * 64; i += 256)
        {
            const int cc = i & 15, o = i >> 4; // U is [oc][c]: consecutive threads read consecutive c
            sU[cc][o] = (oc0 + o < oc && c0 + cc < c) ? __ldg(Uk + (size_t)(oc0 + o) * c + c0 + cc) : 0.f;
        }
        for (int i = threadIdx.x; i < 16 * 64; i += 256)
        {
            const int tt = i & 63, cc = i >> 6;
            sV[cc][tt] = (c0 + cc < c && t0 + tt < T) ? __ldg(Vk + (size_t)(c0 + cc) * T + t0 + tt) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int cc = 0; cc < 16; cc++)
        {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sU[cc][ty * 4 + i], b[i] = sV[cc][tx * 4 + i];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (oc0 + ty * 4 + i < oc && t0 + tx * 4 + j < T) Mk[(size_t)(oc0 + ty * 4 + i) * T + t0 + tx * 4 + j] = acc[i][j];
}

// ---- output transform: Y = A^T m A + bias, activation, cropped to the output ------------------------------------------------------
__global__ void __launch_bounds__(256) wino43_output_kernel(const float* __restrict__ M, const float* __restrict__ bias, float* __restrict__ out, int n, int oc,
                                                            int oh, int ow, int tiles_h, int tiles_w, int activation)
{
    const long long T = (long long)n * tiles_h * tiles_w;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * oc) return;
    const int o = (int)(idx / T);
    const long long t = idx - (long long)o * T;
    const int tw = (int)(t % tiles_w), th = (int)((t / tiles_w) % tiles_h), img = (int)(t / ((long long)tiles_w * tiles_h));
    float m[6][6];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) m[i][j] = __ldg(M + ((size_t)(i * 6 + j) * oc + o) * T + t);
    // A^T (otm): rows {1,1,1,1,1,0} {0,1,-1,2,-2,0} {0,1,1,4,4,0} {0,1,-1,8,-8,1}
    float s[4][6];
#pragma unroll
    for (int j = 0; j < 6; j++)
    {
        const float r0 = m[0][j], r1 = m[1][j], r2 = m[2][j], r3 = m[3][j], r4 = m[4][j], r5 = m[5][j];
        s[0][j] = r0 + r1 + r2 + r3 + r4;
        s[1][j] = r1 - r2 + 2.f * (r3 - r4);
        s[2][j] = r1 + r2 + 4.f * (r3 + r4);
        s[3][j] = r1 - r2 + 8.f * (r3 - r4) + r5;
    }
    const float b = bias ? __ldg(bias + o) : 0.f;
    float* dst = out + ((size_t)img * oc + o) * oh * ow;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const float r0 = s[i][0], r1 = s[i][1], r2 = s[i][2], r3 = s[i][3], r4 = s[i][4], r5 = s[i][5];
        float y[4];
        y[0] = r0 + r1 + r2 + r3 + r4;
        y[1] = r1 - r2 + 2.f * (r3 - r4);
        y[2] = r1 + r2 + 4.f * (r3 + r4);
        y[3] = r1 - r2 + 8.f * (r3 - r4) + r5;
        const int yy = th * 4 + i;
        if (yy < oh)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (tw * 4 + j < ow) dst[(size_t)yy * ow + tw * 4 + j] = act_f32(y[j] + b, activation);
    }
}

size_t wino43_workspace_bytes(int n, int c, int oc, int oh, int ow)
{
    const size_t T = (size_t)n * ((oh + 3) / 4) * ((ow + 3) / 4);
    return sizeof(float) * 36 * ((size_t)oc * c + (size_t)c * T + (size_t)oc * T);
}

cudaError_t launch_conv_winograd43_f32(const float* in, const float* w, const float* bias, float* out, int n, int c, int h, int wd, int oc, int oh, int ow, int ph,
                                       int pw, int activation, float* ws, cudaStream_t st)
{
    const int tiles_h = (oh + 3) / 4, tiles_w = (ow + 3) / 4;
    const long long T = (long long)n * tiles_h * tiles_w;
    if (T * (long long)(c > oc ? c : oc) >= (1ll << 40) || T <= 0) return cudaErrorInvalidValue;
    float* U = ws;
    float* V = U + (size_t)36 * oc * c;
    float* M = V + (size_t)36 * c * T;
    wino43_weight_kernel<<<(oc * c + 255) / 256, 256, 0, st>>>(w, U, oc, c);
    wino43_input_kernel<<<(unsigned)((T * c + 255) / 256), 256, 0, st>>>(in, V, n, c, h, wd, ph, pw, tiles_h, tiles_w);
    const dim3 grid((unsigned)((T + 63) / 64), (unsigned)((oc + 63) / 64), 36);
    wino43_gemm_kernel<<<grid, 256, 0, st>>>(U, V, M, oc, c, T);
    wino43_output_kernel<<<(unsigned)((T * oc + 255) / 256), 256, 0, st>>>(M, bias, out, n, oc, oh, ow, tiles_h, tiles_w, activation);
    return cudaGetLastError();
}

// ---- fp32 depthwise 3x3, stride 1 / 2: one thread per output element ----------------------------------------------------------------
__global__ void __launch_bounds__(256) dw3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                        float* __restrict__ out, int n, int c, int h, int wd, int oh, int ow, int stride, int ph, int pw,
                                                        int activation)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * c * oh * ow;
    if (idx >= total) return;
    const int x = (int)(idx % ow), y = (int)((idx / ow) % oh);
    const long long nc = idx / ((long long)ow * oh);
    const int ch = (int)(nc % c);
    const float* src = in + (size_t)nc * h * wd;
    const float* k = w + (size_t)ch * 9;
    float acc = bias ? __ldg(bias + ch) : 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            const int iy = y * stride - ph + i, ix = x * stride - pw + j;
            if (iy >= 0 && iy < h && ix >= 0 && ix < wd) acc = fmaf(__ldg(src + (size_t)iy * wd + ix), __ldg(k + i * 3 + j), acc);
        }
    out[idx] = act_f32(acc, activation);
}

cudaError_t launch_conv_dw3x3_f32(const float* in, const float* w, const float* bias, float* out, int n, int c, int h, int wd, int oh, int ow, int stride, int ph,
                                  int pw, int activation, cudaStream_t st)
{
    const long long total = (long long)n * c * oh * ow;
    if (total <= 0 || total >= (1ll << 40)) return cudaErrorInvalidValue;
    dw3x3_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, w, bias, out, n, c, h, wd, oh, ow, stride, ph, pw, activation);
    return cudaGetLastError();
}

} // namespace tb200

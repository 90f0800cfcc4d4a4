// engine.cu -- the C ABI of include/tengine_b200.h: context, subgraph planner/executor, weight pre-packing.
//
// One tb200_graph is what the Tengine device glue stores in subgraph->device_graph.  prerun = the analogue of
// the CPU device's create_exec_graph + alloc_exec_graph_mem + prerun_exec_graph (source/device/cpu/
// cpu_device.c:62-95) and of conv_hcl_prerun's weight packing (conv_kernel_x86.c:2137-2209); run = the
// analogue of cpu_device.c:97-221 with every node executing on the GPU.  No CPU compute path exists here.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tengine_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace tb200;

// ---- error reporting -------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_OK(call)                                                                                       \
    do                                                                                                      \
    {                                                                                                       \
        cudaError_t _e = (call);                                                                            \
        if (_e != cudaSuccess) return fail(TB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct tb200_context
{
    int device;
    int num_sms;
    cudaStream_t stream;
};

enum StepKind
{
    K_NCHW2NHWC,
    K_NHWC2NCHW,
    K_CONV_STEM,
    K_STEM_TC,
    K_GATHER_TC,
    K_CONV_DW,
    K_CONV_DIRECT,
    K_GEMM,
    K_IGEMM,
    K_POOL,
    K_POINTWISE,
    K_CONCAT_PART,
    K_UPSAMPLE,
    K_COPY
};
static const char* kStepName[] = {"nchw_to_nhwc", "nhwc_to_nchw", "conv_stem_nchw_dp4a", "conv_stem_nchw_tcgen05", "conv_gather_tcgen05", "conv_dw_direct", "conv_direct_dp4a",
                                  "gemm_i8_tcgen05", "conv_igemm_i8_tcgen05", "pool", "pointwise", "concat_requant", "upsample_nearest", "copy"};

struct Step
{
    int kind;
    int layer; // -1 for layout steps
    const void* in = nullptr;
    const void* in2 = nullptr;
    const void* w = nullptr;
    void* out = nullptr;
    ConvShape cs{};
    PoolShape ps{};
    PointwiseParams pp{};
    EpiParams epi{};
    GemmPlan gemm{};
    const int32_t* btab = nullptr; // uint8 tensor-core kinds: padding-tap corrections
    DwPlan dwp{};
    long long bytes = 0; // pointwise / copy
    // concat / layout
    long long npix = 0;
    int c = 0, cp_in = 0, cp_out = 0, c_off = 0, n = 0, h = 0, w_ = 0, scale = 0;
    int nhwc16 = 0; // K_GATHER_TC: the input is a 16-channel NHWC tensor (else the NCHW network input)
    float s_in = 0, s_out = 0;
    int z_in = 0, z_out = 0;
    bool u8 = false;
};

struct TensorInfo
{
    tb200_tensor_desc d;
    int cp;
    size_t nhwc_bytes, nchw_bytes;
    size_t off; // offset in the activation arena
    uint8_t* dev = nullptr;
    int input_index = -1, output_index = -1;
    bool nhwc_needed = false; // graph inputs: a consumer needs the NHWC copy
    int producer = -1;
};

struct tb200_graph
{
    tb200_context* ctx;
    int flags;
    std::vector<TensorInfo> tensors;
    std::vector<tb200_layer_desc> layers;
    std::vector<const char*> layer_kernel;
    std::vector<Step> steps;
    std::vector<int> input_ids, output_ids;
    std::vector<uint8_t*> in_nchw_dev, out_nchw_dev;
    uint8_t* act_arena = nullptr;
    size_t act_bytes = 0;
    uint8_t* w_arena = nullptr;
    size_t w_bytes = 0;
    int chunks = 1;
    std::vector<std::vector<Step>> chunk_steps;
    std::vector<cudaGraph_t> cu_graphs;
    std::vector<cudaGraphExec_t> cu_execs;
    cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out;
    cudaEvent_t ev_done = nullptr;
    double work_ops = 0, work_bytes = 0;
    int num_launches = 0;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- library / device ------------------------------------------------------------------------------------
extern "C" {

int tb200_abi_version(void) { return TB200_ABI_VERSION; }
const char* tb200_last_error(void) { return g_err; }

int tb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    int good = 0;
    for (int i = 0; i < n; i++)
    {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) good++;
    }
    return good;
}

int tb200_context_create(int cuda_device, tb200_context** out)
{
    if (!out) return fail(TB200_ERR_INVALID, "null out");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    {
        cudaGetLastError();
        return fail(TB200_ERR_NO_DEVICE, "no CUDA device visible: the B200 backend has no CPU fallback");
    }
    if (cuda_device < 0 || cuda_device >= n) return fail(TB200_ERR_INVALID, "cuda device %d out of range (%d)", cuda_device, n);
    int major = 0, minor = 0, sms = 0;
    CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cuda_device));
    CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, cuda_device));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cuda_device));
    if (major != 10) return fail(TB200_ERR_NO_DEVICE, "device %d is sm_%d%d; this backend is built for sm_100a only", cuda_device, major, minor);
    CUDA_OK(cudaSetDevice(cuda_device));
    tb200_context* c = new tb200_context();
    c->device = cuda_device;
    c->num_sms = sms;
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    *out = c;
    return 0;
}

int tb200_context_destroy(tb200_context* ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

void* tb200_context_stream(tb200_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

void* tb200_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void tb200_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int tb200k_cpad(int channels) { return cpad(channels); }

} // extern "C"

// ---- planner ------------------------------------------------------------------------------------------------
static EpiParams make_epi(const tb200_layer_desc& L, const tb200_tensor_desc& tin, const tb200_tensor_desc& tout, bool fc)
{
    EpiParams e{};
    e.in_scale = tin.scale, e.out_scale = tout.scale;
    e.in_zero = tin.zero_point, e.out_zero = tout.zero_point, e.w_zero = L.weight_zero;
    e.activation = L.activation, e.recipe = L.recipe;
    e.is_uint8 = tin.data_type == TB200_DT_UINT8;
    e.fc_rounding = fc ? 1 : 0;
    e.has_bias = L.bias != nullptr;
    e.in_w_scale = e.is_uint8 ? tin.scale * L.weight_scales[0] : 0.f; // fp32 product, as bias_scale in conv_kernel_x86.c:1723
    // fast-path constants (common.cuh requant_fast): clamps of t = f / s_out with activation and saturation folded in
    const float so = tout.scale;
    const float inf = __builtin_inff();
    const int act = fc ? -1 : L.activation;
    float flo = -inf, fhi = inf; // activation clamp in real units
    if (L.recipe == TB200_RECIPE_HCL || fc)
    {
        if (act == 0) flo = 0.f;
        if (act > 0) flo = 0.f, fhi = 6.f;
    }
    else if (act >= 0)
    {
        if (act == 1) flo = -1.f, fhi = 1.f;
        else flo = 0.f, fhi = (act == 6) ? 6.f : inf;
    }
    e.fast_flo = flo, e.fast_fhi = fhi;
    if (e.is_uint8)
    {
        e.fast_lo = (float)(0 - tout.zero_point), e.fast_hi = (float)(255 - tout.zero_point);
        e.fast_r = 1.0f / so;
        // integer-domain clip + zero point + saturation (common.cuh requant_fast8_u8): the activation bounds become the integers
        // the reference's own division and round() give for them; q' = max(min(q + zp - L, H - L), 0), byte = q' + L
        int q_lo = -30000, q_hi = 30000;
        if (flo > -inf) q_lo = (int)roundf(flo / so);
        if (fhi < inf) q_hi = (int)roundf(fhi / so);
        int Lb = q_lo + tout.zero_point, Hb = q_hi + tout.zero_point;
        Lb = Lb < 0 ? 0 : (Lb > 255 ? 255 : Lb), Hb = Hb > 255 ? 255 : (Hb < 0 ? 0 : Hb);
        const uint32_t add = (uint32_t)(tout.zero_point - Lb) & 0xffffu, mx = (uint32_t)(Hb - Lb) & 0xffffu;
        e.q_add2 = add | (add << 16), e.q_max2 = mx | (mx << 16);
        e.q_byte_add = ((uint32_t)Lb & 0xffu) * 0x01010101u;
    }
    else
    {
        e.fast_lo = -127.f, e.fast_hi = 127.f;
        if (flo > -inf) e.fast_lo = fmaxf(-127.f, flo / so); // fl(x / s_out): the same rounded quotient the reference forms
        if (fhi < inf) e.fast_hi = fminf(127.f, fhi / so);
        e.fast_r = 0.f;
        // integer-domain clamp (common.cuh requant_fast4_i8): [q_lo, q_hi] = what the reference's roundf gives for the
        // clipped bounds; q' = max(min(q - q_lo, q_hi - q_lo), 0), bytes = q' + q_lo
        const int q_lo = (int)roundf(e.fast_lo), q_hi = (int)roundf(e.fast_hi);
        const uint32_t add = (uint32_t)(-q_lo) & 0xffffu, mx = (uint32_t)(q_hi - q_lo) & 0xffffu;
        e.q_add2 = add | (add << 16), e.q_max2 = mx | (mx << 16);
        e.q_byte_add = ((uint32_t)q_lo & 0xffu) * 0x01010101u;
    }
    e.fast_ok = (so > 1e-30f && so < 1e30f && tin.scale > 1e-30f && tin.scale < 1e30f) ? 1 : 0;
    return e;
}

// bytes of the packed B operand of the tensor-core paths: [n_tiles][block_n (+16 for uint8: the ones-row)][K]
static size_t gemm_weight_bytes(int ocp, int k, bool u8)
{
    const int bn = gemm_block_n(ocp, u8), nt = (ocp + bn - 1) / bn;
    return (size_t)nt * (bn + (u8 ? 16 : 0)) * k;
}

struct WeightBlob
{
    size_t w_off, bias_off, scale_off, fast_off, btab_off, w_size;
};

static int run_step(tb200_graph* g, const Step& s, cudaStream_t st)
{
    cudaError_t err = cudaSuccess;
    switch (s.kind)
    {
    case K_NCHW2NHWC: err = launch_nchw_to_nhwc(s.in, s.out, s.n, s.c, s.h, s.w_, st); break;
    case K_NHWC2NCHW: err = launch_nhwc_to_nchw(s.in, s.out, s.n, s.c, s.h, s.w_, st); break;
    case K_CONV_STEM: err = launch_conv_stem(s.in, s.w, s.out, s.cs, s.epi, st); break;
    case K_STEM_TC: err = launch_stem_tc(s.dwp, s.in, s.w, s.out, s.cs, s.epi, st); break;
    case K_GATHER_TC: err = launch_conv_gather_tc(s.in, s.w, s.out, s.cs, s.epi, s.nhwc16, st); break;
    case K_CONV_DW:
        err = s.dwp.valid ? launch_conv_dw_tma(s.dwp, s.w, s.out, s.cs, s.epi, st) : launch_conv_dw(s.in, s.w, s.out, s.cs, s.epi, st);
        break;
    case K_CONV_DIRECT: err = launch_conv_direct(s.in, s.w, s.out, s.cs, s.epi, st); break;
    case K_GEMM:
    case K_IGEMM: err = launch_gemm_i8(s.gemm, s.epi, s.btab, g->ctx->num_sms, st); break;
    case K_POOL: err = launch_pool(s.in, s.out, s.ps, s.u8, st); break;
    case K_POINTWISE: err = launch_pointwise(s.in, s.in2, s.out, s.bytes, s.pp, s.u8, st); break;
    case K_CONCAT_PART:
        err = launch_concat_part(s.in, s.out, s.npix, s.c, s.cp_in, s.cp_out, s.c_off, s.s_in, s.z_in, s.s_out, s.z_out, s.u8, st);
        break;
    case K_UPSAMPLE: err = launch_upsample(s.in, s.out, s.n, s.h, s.w_, s.cp_in, s.scale, st); break;
    case K_COPY: err = cudaMemcpyAsync(s.out, s.in, (size_t)s.bytes, cudaMemcpyDeviceToDevice, st); break;
    }
    if (err != cudaSuccess) return fail(TB200_ERR_CUDA, "launch of %s (layer %d) failed: %s", kStepName[s.kind], s.layer, cudaGetErrorString(err));
    return 0;
}

static void destroy_graph(tb200_graph* g)
{
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    for (auto e : g->cu_execs) if (e) cudaGraphExecDestroy(e);
    for (auto c : g->cu_graphs) if (c) cudaGraphDestroy(c);
    for (auto e : g->ev_in) cudaEventDestroy(e);
    for (auto e : g->ev_out) cudaEventDestroy(e);
    if (g->ev_done) cudaEventDestroy(g->ev_done);
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    if (g->d2h_stream) cudaStreamDestroy(g->d2h_stream);
    for (auto p : g->in_nchw_dev) cudaFree(p);
    for (auto p : g->out_nchw_dev) cudaFree(p);
    cudaFree(g->act_arena);
    cudaFree(g->w_arena);
    delete g;
}

extern "C" int tb200_graph_prerun(tb200_context* ctx, const tb200_tensor_desc* tensors, int num_tensors,
                                  const tb200_layer_desc* layers, int num_layers, const int32_t* input_ids, int num_inputs,
                                  const int32_t* output_ids, int num_outputs, int flags, tb200_graph** out)
{
    if (!ctx || !tensors || !layers || !out || num_tensors <= 0 || num_layers <= 0) return fail(TB200_ERR_INVALID, "bad arguments");
    CUDA_OK(cudaSetDevice(ctx->device));
    tb200_graph* g = new tb200_graph();
    g->ctx = ctx, g->flags = flags;
    auto bail = [&](int rc) { destroy_graph(g); return rc; };

    // ---- tensors ----
    g->tensors.resize(num_tensors);
    size_t act = 0;
    for (int i = 0; i < num_tensors; i++)
    {
        TensorInfo& t = g->tensors[i];
        t.d = tensors[i];
        if (t.d.data_type != TB200_DT_INT8 && t.d.data_type != TB200_DT_UINT8)
            return bail(fail(TB200_ERR_UNSUPPORTED, "tensor %d: data type %d not supported by the B200 backend", i, t.d.data_type));
        for (int k = 0; k < 4; k++)
            if (t.d.dims[k] <= 0) return bail(fail(TB200_ERR_INVALID, "tensor %d: dim %d is %d", i, k, t.d.dims[k]));
        t.cp = cpad(t.d.dims[1]);
        t.nhwc_bytes = (size_t)t.d.dims[0] * t.d.dims[2] * t.d.dims[3] * t.cp;
        t.nchw_bytes = (size_t)t.d.dims[0] * t.d.dims[1] * t.d.dims[2] * t.d.dims[3];
        t.off = act;
        act += align_up(t.nhwc_bytes, 1024);
    }
    for (int i = 0; i < num_inputs; i++)
    {
        if (input_ids[i] < 0 || input_ids[i] >= num_tensors) return bail(fail(TB200_ERR_INVALID, "input id out of range"));
        g->tensors[input_ids[i]].input_index = i;
        g->input_ids.push_back(input_ids[i]);
    }
    for (int i = 0; i < num_outputs; i++)
    {
        if (output_ids[i] < 0 || output_ids[i] >= num_tensors) return bail(fail(TB200_ERR_INVALID, "output id out of range"));
        g->tensors[output_ids[i]].output_index = i;
        g->output_ids.push_back(output_ids[i]);
    }
    g->layers.assign(layers, layers + num_layers);
    g->layer_kernel.assign(num_layers, "");

    // ---- validate layers, decide kernels, size the weight arena ----
    std::vector<int> kind(num_layers, -1);
    std::vector<WeightBlob> blobs(num_layers);
    std::vector<int> fuse_bias(num_layers, 0);
    size_t wtotal = 0;
    const bool no_tc = (flags & TB200_PRERUN_NO_TENSORCORE) != 0;
    for (int li = 0; li < num_layers; li++)
    {
        const tb200_layer_desc& L = layers[li];
        if (L.num_inputs < 1 || L.num_inputs > 4) return bail(fail(TB200_ERR_INVALID, "layer %d: num_inputs %d", li, L.num_inputs));
        for (int k = 0; k < L.num_inputs; k++)
            if (L.inputs[k] < 0 || L.inputs[k] >= num_tensors) return bail(fail(TB200_ERR_INVALID, "layer %d: input id", li));
        if (L.output < 0 || L.output >= num_tensors) return bail(fail(TB200_ERR_INVALID, "layer %d: output id", li));
        TensorInfo& tin = g->tensors[L.inputs[0]];
        TensorInfo& tout = g->tensors[L.output];
        tout.producer = li;
        if (tin.d.data_type != tout.d.data_type) return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: mixed data types", li));
        const bool u8 = tin.d.data_type == TB200_DT_UINT8;
        const int C = tin.d.dims[1], H = tin.d.dims[2], W = tin.d.dims[3], OC = tout.d.dims[1];
        if (L.op == TB200_OP_CONV)
        {
            if (!L.weight || !L.weight_scales) return bail(fail(TB200_ERR_INVALID, "layer %d: conv without weight/scales", li));
            if (L.group < 1 || C % L.group || OC % L.group) return bail(fail(TB200_ERR_INVALID, "layer %d: bad group %d", li, L.group));
            const int oh = (H + L.pad_h0 + L.pad_h1 - (L.dilation_h * (L.kernel_h - 1) + 1)) / L.stride_h + 1;
            const int ow = (W + L.pad_w0 + L.pad_w1 - (L.dilation_w * (L.kernel_w - 1) + 1)) / L.stride_w + 1;
            if (oh != tout.d.dims[2] || ow != tout.d.dims[3] || tout.d.dims[0] != tin.d.dims[0])
                return bail(fail(TB200_ERR_INVALID, "layer %d: conv output shape %dx%d does not match descriptor %dx%d", li, oh, ow, tout.d.dims[2], tout.d.dims[3]));
            const int cg = C / L.group;
            size_t wsize;
            if (tin.input_index >= 0 && C <= 3 && L.group == 1 && !u8 && !no_tc && L.kernel_h == 3 && L.kernel_w == 3 && L.dilation_h == 1 &&
                L.dilation_w == 1 && L.stride_h == L.stride_w && tout.cp <= 256 && !getenv("TB200_NO_STEM_TC"))
                kind[li] = K_STEM_TC, wsize = (size_t)tout.cp * 32; // one 32-byte UMMA k-step per output channel
            else if (tin.input_index >= 0 && C <= 3 && L.group == 1 && !no_tc && L.kernel_h == L.kernel_w && (L.kernel_h == 7 || (u8 && L.kernel_h == 3)) &&
                     L.dilation_h == 1 && L.dilation_w == 1 && L.stride_h == L.stride_w && tout.cp <= 256 && !getenv("TB200_NO_GATHER_TC"))
                // uint8 3x3 stems and 7x7 stems (ResNet): threads gather, taps outside the image = zero point; K padded to 32*ks
                kind[li] = K_GATHER_TC, wsize = (size_t)tout.cp * 32 * ((C * L.kernel_h * L.kernel_w + 31) / 32);
            else if (tin.input_index >= 0 && C <= 4 && L.group == 1)
                kind[li] = K_CONV_STEM, wsize = (size_t)tout.cp * L.kernel_h * L.kernel_w * 4;
            else if (!no_tc && L.group == 1 && tin.cp == 16 && L.kernel_h == 3 && L.kernel_w == 3 && L.dilation_h == 1 && L.dilation_w == 1 &&
                     L.stride_h == L.stride_w && tout.cp <= 256 && !getenv("TB200_NO_GATHER_TC"))
                kind[li] = K_GATHER_TC, wsize = (size_t)tout.cp * 160; // 16-channel input: nine 16-byte taps per pixel, five k-steps
            else if (L.group == C && OC == C && C > 1)
                kind[li] = K_CONV_DW, wsize = (size_t)L.kernel_h * L.kernel_w * tin.cp;
            else if (!no_tc && L.group == 1 && L.kernel_h == 1 && L.kernel_w == 1 && L.stride_h == 1 && L.stride_w == 1 &&
                     !L.pad_h0 && !L.pad_h1 && !L.pad_w0 && !L.pad_w1 && !(u8 && getenv("TB200_NO_U8_TC")))
                kind[li] = K_GEMM, wsize = gemm_weight_bytes(tout.cp, tin.cp, u8);
            else if (!no_tc && L.group == 1 && L.dilation_h == 1 && L.dilation_w == 1 && L.stride_h == L.stride_w &&
                     (L.stride_h == 1 || L.stride_h == 2) && (L.kernel_h * L.kernel_w == 1 || tin.cp % 32 == 0) && tout.d.dims[3] <= 4096 &&
                     L.kernel_h * L.kernel_w <= 64 && !getenv("TB200_NO_IGEMM") && !(u8 && getenv("TB200_NO_U8_TC")))
                kind[li] = K_IGEMM, wsize = gemm_weight_bytes(tout.cp, L.kernel_h * L.kernel_w * tin.cp, u8);
            else
            {
                if (L.group > 1 && ((cg % 4) || ((OC / L.group) % 4)))
                    return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: grouped conv needs channels per group %% 4 == 0", li));
                kind[li] = K_CONV_DIRECT;
                wsize = (size_t)tout.cp * L.kernel_h * L.kernel_w * (L.group == 1 ? tin.cp : cg);
            }
            blobs[li].w_size = wsize;
        }
        else if (L.op == TB200_OP_FC)
        {
            if (!L.weight || !L.weight_scales) return bail(fail(TB200_ERR_INVALID, "layer %d: fc without weight/scales", li));
            if (no_tc || (u8 && getenv("TB200_NO_U8_TC")))
                kind[li] = K_CONV_DIRECT, blobs[li].w_size = (size_t)tout.cp * H * W * tin.cp; // FC == conv with kernel HxW over the whole input
            else
                kind[li] = K_GEMM, blobs[li].w_size = gemm_weight_bytes(tout.cp, H * W * tin.cp, u8);
        }
        else if (L.op == TB200_OP_POOL)
            kind[li] = K_POOL;
        else if (L.op == TB200_OP_RELU)
            kind[li] = K_POINTWISE;
        else if (L.op == TB200_OP_ELTWISE)
        {
            if (L.num_inputs != 2 || (L.elt_type != TB200_ELT_SUM && L.elt_type != TB200_ELT_PROD))
                return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: eltwise type %d / %d inputs", li, L.elt_type, L.num_inputs));
            kind[li] = K_POINTWISE;
        }
        else if (L.op == TB200_OP_CONCAT)
        {
            if (L.axis != 1) return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: concat axis %d", li, L.axis));
            kind[li] = (L.num_inputs == 1) ? K_COPY : K_CONCAT_PART; // concat_kernel_ref_int8.c:45-56: one input is a plain copy
        }
        else if (L.op == TB200_OP_UPSAMPLE)
            kind[li] = K_UPSAMPLE;
        else if (L.op == TB200_OP_IDENTITY)
            kind[li] = K_COPY;
        else
            return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: op %d", li, L.op));

        if (L.op == TB200_OP_CONV || L.op == TB200_OP_FC)
        {
            blobs[li].w_off = wtotal;
            wtotal += align_up(blobs[li].w_size, 1024);
            blobs[li].bias_off = wtotal;
            wtotal += align_up((size_t)tout.cp * 4, 256);
            blobs[li].scale_off = wtotal;
            wtotal += align_up((size_t)tout.cp * 4, 256);
            blobs[li].fast_off = wtotal;
            wtotal += align_up((size_t)tout.cp * 8, 256);
            blobs[li].btab_off = wtotal;
            if (u8 && (kind[li] == K_GEMM || kind[li] == K_IGEMM))
                wtotal += align_up((size_t)(L.op == TB200_OP_FC ? 1 : L.kernel_h * L.kernel_w) * tout.cp * 4, 256);
        }
        // which graph inputs need an NHWC copy (everything except a stem conv reads NHWC)
        for (int k = 0; k < L.num_inputs; k++)
            if (g->tensors[L.inputs[k]].input_index >= 0 && kind[li] != K_CONV_STEM && kind[li] != K_STEM_TC && !(kind[li] == K_GATHER_TC && g->tensors[L.inputs[k]].d.dims[1] <= 3)) g->tensors[L.inputs[k]].nhwc_needed = true;
    }
    for (int id : g->output_ids)
        if (g->tensors[id].input_index >= 0) g->tensors[id].nhwc_needed = true;

    // ---- allocate ----
    g->act_bytes = act;
    CUDA_OK(cudaMalloc(&g->act_arena, act ? act : 1024));
    CUDA_OK(cudaMemsetAsync(g->act_arena, 0, act ? act : 1024, ctx->stream));
    for (auto& t : g->tensors) t.dev = g->act_arena + t.off;
    g->w_bytes = wtotal ? wtotal : 1024;
    CUDA_OK(cudaMalloc(&g->w_arena, g->w_bytes));
    for (int id : g->input_ids)
    {
        uint8_t* p = nullptr;
        CUDA_OK(cudaMalloc(&p, g->tensors[id].nchw_bytes));
        g->in_nchw_dev.push_back(p);
    }
    for (int id : g->output_ids)
    {
        uint8_t* p = nullptr;
        CUDA_OK(cudaMalloc(&p, g->tensors[id].nchw_bytes));
        g->out_nchw_dev.push_back(p);
    }

    // ---- per layer: may the fast epilogue fold the bias add into its FMA?  (common.cuh requant_fast_bits<.., FUSE>)
    //      Decided from the descriptors alone so that every rank of a sharded job reaches the same arena format. ----
    for (int li = 0; li < num_layers; li++)
    {
        const tb200_layer_desc& L = layers[li];
        if (L.op != TB200_OP_CONV) continue;
        const TensorInfo& tin = g->tensors[L.inputs[0]];
        const TensorInfo& tout = g->tensors[L.output];
        if (tin.d.data_type == TB200_DT_UINT8 || getenv("TB200_NO_FUSE_BIAS")) continue;
        bool fuse = true;
        for (int o = 0; o < tout.d.dims[1]; o++)
        {
            const double bm = (double)(L.bias ? L.bias[o] : 0) * (double)tin.d.scale * (double)L.weight_scales[o] / (double)tout.d.scale;
            if (!(bm >= -100.0 && bm <= 100.0)) fuse = false;
        }
        fuse_bias[li] = fuse ? 1 : 0;
    }
    // ---- int8 fast epilogue rounds BEFORE it clamps (integer-domain clamp on 16-bit lanes, common.cuh): prove from the
    //      weights that no reachable accumulator can leave the 16-bit range after scaling, else use the exact path ----
    std::vector<int> fast_int_ok(num_layers, 1);
    for (int li = 0; li < num_layers; li++)
    {
        const tb200_layer_desc& L = layers[li];
        if (L.op != TB200_OP_CONV && L.op != TB200_OP_FC) continue;
        const TensorInfo& tin = g->tensors[L.inputs[0]];
        const TensorInfo& tout = g->tensors[L.output];
        const bool u8b = tin.d.data_type == TB200_DT_UINT8;
        const int OC = tout.d.dims[1];
        const size_t kk = L.op == TB200_OP_FC ? (size_t)tin.d.dims[1] * tin.d.dims[2] * tin.d.dims[3]
                                              : (size_t)(tin.d.dims[1] / L.group) * L.kernel_h * L.kernel_w;
        const int8_t* wsrc = (const int8_t*)L.weight;
        for (int o = 0; o < OC; o++)
        {
            double sumabs = 0;
            if (u8b)
                for (size_t k = 0; k < kk; k++) sumabs += abs((int)((const uint8_t*)L.weight)[(size_t)o * kk + k] - L.weight_zero);
            else
                for (size_t k = 0; k < kk; k++) sumabs += abs((int)wsrc[(size_t)o * kk + k]);
            const double M = (double)tin.d.scale * (double)L.weight_scales[u8b ? 0 : o] / (double)tout.d.scale;
            const double bound = ((u8b ? 255.0 : 128.0) * sumabs + fabs((double)(L.bias ? L.bias[o] : 0))) * fabs(M) + 256.0;
            if (!(bound < 32000.0)) fast_int_ok[li] = 0;
        }
    }

    // ---- pack weights into a host image of the arena, one H2D copy ----
    if (!(flags & TB200_PRERUN_NO_WEIGHTS))
    {
        std::vector<uint8_t> img(g->w_bytes, 0);
        for (int li = 0; li < num_layers; li++)
        {
            const tb200_layer_desc& L = layers[li];
            if (L.op != TB200_OP_CONV && L.op != TB200_OP_FC) continue;
            const TensorInfo& tin = g->tensors[L.inputs[0]];
            const TensorInfo& tout = g->tensors[L.output];
            const bool u8 = tin.d.data_type == TB200_DT_UINT8;
            const int C = tin.d.dims[1], H = tin.d.dims[2], W = tin.d.dims[3], OC = tout.d.dims[1];
            const uint8_t* src = (const uint8_t*)L.weight;
            uint8_t* dst = img.data() + blobs[li].w_off;
            const bool tc = kind[li] == K_GEMM || kind[li] == K_IGEMM;
            // row of output channel o in the packed B operand (uint8 tensor-core tiles carry 16 extra rows each)
            const int bn = tc ? gemm_block_n(tout.cp, u8) : tout.cp, bnx = bn + ((tc && u8) ? 16 : 0);
            auto brow = [&](int o) -> size_t { return (size_t)(o / bn) * bnx + (o % bn); };
            size_t krow = 0; // K extent of one packed row
            if (L.op == TB200_OP_FC)
            {
                // [OC][C*H*W] (NCHW flatten, fc_ref.c:313-359) -> [row][H][W][Cp]
                krow = (size_t)H * W * tin.cp;
                for (int o = 0; o < OC; o++)
                    for (int c = 0; c < C; c++)
                        for (int p = 0; p < H * W; p++) dst[(brow(o) * H * W + p) * tin.cp + c] = src[((size_t)o * C + c) * H * W + p];
            }
            else
            {
                const int KH = L.kernel_h, KW = L.kernel_w, cg = C / L.group;
                if (kind[li] == K_CONV_DW)
                {
                    for (int c = 0; c < C; c++)
                        for (int t = 0; t < KH * KW; t++) dst[(size_t)t * tin.cp + c] = src[(size_t)c * KH * KW + t];
                }
                else if (kind[li] == K_GATHER_TC && !(tin.input_index >= 0 && C <= 3))
                {
                    // 16-channel NHWC input: k = (kh*3 + kw)*16 + c, 160-byte rows (the last 16 bytes stay 0)
                    for (int o = 0; o < OC; o++)
                        for (int c = 0; c < C; c++)
                            for (int t = 0; t < 9; t++) dst[(size_t)o * 160 + t * 16 + c] = src[((size_t)o * C + c) * 9 + t];
                }
                else if (kind[li] == K_STEM_TC || kind[li] == K_GATHER_TC)
                {
                    // [OC][C][KH][KW] is already k = (c*KH + kh)*KW + kw order: one zero-padded row of 32*ks bytes per channel
                    const size_t kk = (size_t)C * KH * KW, kp = ((kk + 31) / 32) * 32;
                    for (int o = 0; o < OC; o++) memcpy(dst + (size_t)o * kp, src + (size_t)o * kk, kk);
                }
                else
                {
                    const int cgp = kind[li] == K_CONV_STEM ? 4 : (L.group == 1 ? tin.cp : cg);
                    krow = (size_t)KH * KW * cgp;
                    for (int o = 0; o < OC; o++)
                        for (int c = 0; c < cg; c++)
                            for (int t = 0; t < KH * KW; t++)
                                dst[(brow(o) * KH * KW + t) * cgp + c] = src[((size_t)o * cg + c) * KH * KW + t];
                }
            }
            std::vector<int64_t> wsum_tot(tout.cp, 0); // uint8 tensor-core kinds: sum of the raw weight bytes per channel
            if (tc && u8)
            {
                // ones-row of every N tile: accumulator column block_n becomes sum_k x of the pixel
                const int nt = (tout.cp + bn - 1) / bn;
                for (int t = 0; t < nt; t++) memset(dst + ((size_t)t * bnx + bn) * krow, 1, krow);
                // pad K positions (channels >= C) of the ones-row must not count: x is 0 there anyway (pad lanes hold 0)
                const int taps = L.op == TB200_OP_FC ? 1 : L.kernel_h * L.kernel_w;
                const int kk = L.op == TB200_OP_FC ? C * H * W : C; // real K elements per tap
                int32_t* bt = (int32_t*)(img.data() + blobs[li].btab_off);
                for (int o = 0; o < OC; o++)
                    for (int t = 0; t < taps; t++)
                    {
                        int64_t sw = 0;
                        if (L.op == TB200_OP_FC)
                            for (int k = 0; k < kk; k++) sw += src[(size_t)o * kk + k];
                        else
                            for (int c = 0; c < C; c++) sw += src[((size_t)o * C + c) * taps + t];
                        wsum_tot[o] += sw;
                        // what a tap that falls into the padding must give back: zx*(sum_c w - Cin*zw)
                        bt[(size_t)t * tout.cp + o] = (int32_t)((int64_t)tin.d.zero_point * (sw - (int64_t)kk * L.weight_zero));
                    }
            }
            if (kind[li] == K_GATHER_TC && u8)
            {
                const int kk = C * L.kernel_h * L.kernel_w;
                for (int o = 0; o < OC; o++)
                    for (int k = 0; k < kk; k++) wsum_tot[o] += src[(size_t)o * kk + k];
            }
            int32_t* b = (int32_t*)(img.data() + blobs[li].bias_off);
            float* sc = (float*)(img.data() + blobs[li].scale_off);
            float* fm = (float*)(img.data() + blobs[li].fast_off); // float2 per channel: (multiplier | bias term, bias bits)
            const bool fc = L.op == TB200_OP_FC;
            const bool fuse = fuse_bias[li] != 0;
            for (int o = 0; o < tout.cp; o++)
            {
                b[o] = (L.bias && o < OC) ? L.bias[o] : 0;
                sc[o] = u8 ? L.weight_scales[0] : (o < OC ? L.weight_scales[o] : 1.f);
                if (u8)
                {
                    // the bias term in real units, rounded exactly as the reference rounds it
                    const float S = tin.d.scale * L.weight_scales[0];
                    fm[2 * o] = (o >= OC) ? 0.f : ((L.recipe == TB200_RECIPE_HCL || fc) ? (float)b[o] * S : ((float)b[o] * tin.d.scale) * L.weight_scales[0]);
                    // tensor-core kinds: corr[oc] = -zx*sum_k w + K*zx*zw (all taps in bounds), carried in the .y lane
                    int32_t corr = 0;
                    if ((tc || kind[li] == K_GATHER_TC) && o < OC)
                    {
                        const int64_t kreal = fc ? (int64_t)C * H * W : (int64_t)C * L.kernel_h * L.kernel_w;
                        corr = (int32_t)(-(int64_t)tin.d.zero_point * wsum_tot[o] + kreal * tin.d.zero_point * L.weight_zero);
                    }
                    memcpy(&fm[2 * o + 1], &corr, 4);
                }
                else
                {
                    // channel pairs interleaved (common.cuh FastPar4): { M[2k], M[2k+1], y[2k], y[2k+1] }
                    float* fmM = fm + (size_t)(o >> 1) * 4 + (o & 1);
                    float* fmY = fmM + 2;
                    if (o >= OC) *fmM = 0.f;
                    else if (fc) *fmM = (tin.d.scale * sc[o]) / tout.d.scale; // fc_ref.c:225, the reference's own requant scale
                    else *fmM = (float)((double)tin.d.scale * (double)sc[o] / (double)tout.d.scale);
                    if (fuse) *fmY = (o >= OC) ? 0.f : (float)((double)b[o] * (double)*fmM); // fl(bias*M)
                    else memcpy(fmY, &b[o], 4);
                }
            }
        }
        CUDA_OK(cudaMemcpyAsync(g->w_arena, img.data(), g->w_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_OK(cudaStreamSynchronize(ctx->stream));
    }

    // ---- build the launch sequence: the batch is cut into K equal chunks (images are independent units) so that `run`
    //      can overlap the H2D copy of chunk i+1 with the kernels of chunk i; each chunk owns a slice of every tensor ----
    int Ntot = g->tensors[0].d.dims[0];
    bool same_batch = true;
    for (auto& t : g->tensors) same_batch &= (t.d.dims[0] == Ntot);
    int K = 1;
    if (same_batch && !(flags & TB200_PRERUN_NO_GRAPH))
    {
        K = Ntot >= 32 ? 2 : 1;
        if (const char* ev = getenv("TB200_PIPELINE_CHUNKS")) K = atoi(ev) > 0 ? atoi(ev) : K;
        while (K > 1 && Ntot % K) K--;
    }
    const int Kpipe = K;
    g->chunks = Kpipe;
    g->chunk_steps.resize(Kpipe > 1 ? Kpipe : 0);
    // builds the launch sequence of chunk `ck` of `K` (K == 1: the whole batch) into `steps`
    auto build_steps = [&](const int K, const int ck, std::vector<Step>& steps, const bool count_work) -> int
    {
        const int nb = Ntot / K;
        auto tdev = [&](const TensorInfo& t) -> uint8_t* { return t.dev + (size_t)ck * (t.nhwc_bytes / K); };
        for (size_t i = 0; i < g->input_ids.size(); i++)
        {
            TensorInfo& t = g->tensors[g->input_ids[i]];
            if (!t.nhwc_needed) continue;
            Step s;
            s.kind = K_NCHW2NHWC, s.layer = -1, s.in = g->in_nchw_dev[i] + ck * (t.nchw_bytes / K), s.out = tdev(t);
            s.n = nb, s.c = t.d.dims[1], s.h = t.d.dims[2], s.w_ = t.d.dims[3];
            steps.push_back(s);
        }
        for (int li = 0; li < num_layers; li++)
        {
            const tb200_layer_desc& L = layers[li];
            TensorInfo& tin = g->tensors[L.inputs[0]];
            TensorInfo& tout = g->tensors[L.output];
            const bool u8 = tin.d.data_type == TB200_DT_UINT8;
            Step s;
            s.kind = kind[li], s.layer = li, s.u8 = u8;
            s.in = tdev(tin), s.out = tdev(tout);
            const int N = nb, C = tin.d.dims[1], H = tin.d.dims[2], W = tin.d.dims[3];
            const int OC = tout.d.dims[1], OH = tout.d.dims[2], OW = tout.d.dims[3];
            if (L.op == TB200_OP_CONV || L.op == TB200_OP_FC)
            {
                const bool fc = L.op == TB200_OP_FC;
                s.w = g->w_arena + blobs[li].w_off;
                s.epi = make_epi(L, tin.d, tout.d, fc);
                s.epi.bias = (const int32_t*)(g->w_arena + blobs[li].bias_off);
                s.epi.w_scale = (const float*)(g->w_arena + blobs[li].scale_off);
                s.epi.fast_par = (const float2*)(g->w_arena + blobs[li].fast_off);
                s.btab = (const int32_t*)(g->w_arena + blobs[li].btab_off);
                s.epi.fuse_bias = fuse_bias[li];
                if (!fast_int_ok[li]) s.epi.fast_ok = 0;
                ConvShape& cs = s.cs;
                cs.n = N, cs.h = H, cs.w = W, cs.c = C, cs.cp = tin.cp, cs.oh = OH, cs.ow = OW, cs.oc = OC, cs.ocp = tout.cp;
                if (fc)
                {
                    cs.kh = H, cs.kw = W, cs.sh = cs.sw = 1, cs.ph0 = cs.pw0 = 0, cs.dh = cs.dw = 1, cs.group = 1;
                    cs.cg = C, cs.cgp = tin.cp;
                    if (count_work)
                    {
                        g->work_ops += 2.0 * Ntot * OC * C * H * W;
                        g->work_bytes += (double)tin.nchw_bytes + tout.nchw_bytes + (double)OC * C * H * W + (L.bias ? 4.0 * OC : 0);
                    }
                }
                else
                {
                    cs.kh = L.kernel_h, cs.kw = L.kernel_w, cs.sh = L.stride_h, cs.sw = L.stride_w, cs.ph0 = L.pad_h0, cs.pw0 = L.pad_w0;
                    cs.dh = L.dilation_h, cs.dw = L.dilation_w, cs.group = L.group;
                    cs.cg = C / L.group, cs.cgp = (L.group == 1) ? tin.cp : cs.cg;
                    const double k = (double)cs.cg * cs.kh * cs.kw;
                    if (count_work)
                    {
                        g->work_ops += 2.0 * (double)tout.nchw_bytes * k;
                        g->work_bytes += (double)tin.nchw_bytes + tout.nchw_bytes + (double)OC * k + (L.bias ? 4.0 * OC : 0);
                    }
                }
                if (s.kind == K_CONV_STEM || s.kind == K_STEM_TC || (s.kind == K_GATHER_TC && tin.input_index >= 0 && C <= 3))
                    s.in = g->in_nchw_dev[tin.input_index] + ck * (tin.nchw_bytes / K);
                s.nhwc16 = (s.kind == K_GATHER_TC && !(tin.input_index >= 0 && C <= 3)) ? 1 : 0;
                if (s.kind == K_STEM_TC) stem_plan_create(&s.dwp, s.in, s.cs); // falls back to the global-memory gather
                if (s.kind == K_CONV_DW && !(flags & TB200_PRERUN_NO_TENSORCORE)) dw_plan_create(&s.dwp, s.in, s.cs, s.epi); // falls back when not applicable
                if (s.kind == K_IGEMM)
                {
                    int rc = gemm_plan_create_conv(&s.gemm, s.in, s.w, s.out, s.cs, u8 ? 1 : 0);
                    if (rc) return bail(fail(rc, "layer %d: implicit-GEMM plan failed", li));
                }
                if (s.kind == K_GEMM)
                {
                    const long long m = fc ? N : (long long)N * H * W;
                    const int kdim = fc ? H * W * tin.cp : tin.cp;
                    int rc = gemm_plan_create(&s.gemm, s.in, kdim, s.w, s.out, m, kdim, OC, tout.cp, tout.cp, 0, u8 ? 1 : 0);
                    if (rc) return bail(fail(rc, "layer %d: TMA descriptor creation failed (m=%lld k=%d oc=%d)", li, m, kdim, OC));
                }
            }
            else if (L.op == TB200_OP_POOL)
            {
                PoolShape& p = s.ps;
                p.n = N, p.h = H, p.w = W, p.c = C, p.cp = tin.cp, p.oh = OH, p.ow = OW;
                p.kh = L.kernel_h, p.kw = L.kernel_w, p.sh = L.stride_h, p.sw = L.stride_w, p.ph0 = L.pad_h0, p.pw0 = L.pad_w0;
                p.method = L.pool_method, p.caffe_flavor = L.caffe_flavor;
                p.in_scale = tin.d.scale, p.out_scale = tout.d.scale, p.in_zero = tin.d.zero_point, p.out_zero = tout.d.zero_point;
                if (L.pool_global) p.kh = H, p.kw = W, p.sh = p.sw = 1, p.ph0 = p.pw0 = 0;
                if (tout.d.dims[1] != C) return bail(fail(TB200_ERR_INVALID, "layer %d: pool channel mismatch", li));
            }
            else if (L.op == TB200_OP_RELU || L.op == TB200_OP_ELTWISE)
            {
                PointwiseParams& p = s.pp;
                p.c = C, p.cp = tin.cp;
                p.scale0 = tin.d.scale, p.zero0 = tin.d.zero_point, p.out_scale = tout.d.scale, p.out_zero = tout.d.zero_point;
                p.negative_slope = L.negative_slope;
                if (L.op == TB200_OP_RELU)
                    p.mode = 0, p.scale1 = 0, p.zero1 = 0;
                else
                {
                    const TensorInfo& t1 = g->tensors[L.inputs[1]];
                    if (t1.nhwc_bytes != tin.nhwc_bytes) return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: eltwise broadcast", li));
                    p.mode = L.elt_type == TB200_ELT_SUM ? 1 : 2;
                    p.scale1 = t1.d.scale, p.zero1 = t1.d.zero_point;
                    s.in2 = tdev(t1);
                }
                s.bytes = (long long)(tin.nhwc_bytes / K);
                if (tout.nhwc_bytes != tin.nhwc_bytes) return bail(fail(TB200_ERR_INVALID, "layer %d: pointwise shape mismatch", li));
            }
            else if (L.op == TB200_OP_CONCAT && L.num_inputs > 1)
            {
                int coff = 0;
                for (int k = 0; k < L.num_inputs; k++)
                {
                    const TensorInfo& tk = g->tensors[L.inputs[k]];
                    Step p = s;
                    p.in = tdev(tk);
                    p.npix = (long long)N * H * W, p.c = tk.d.dims[1], p.cp_in = tk.cp, p.cp_out = tout.cp, p.c_off = coff;
                    p.s_in = tk.d.scale, p.z_in = tk.d.zero_point, p.s_out = tout.d.scale, p.z_out = tout.d.zero_point;
                    coff += tk.d.dims[1];
                    if (k + 1 < L.num_inputs) steps.push_back(p);
                    else s = p;
                }
                if (coff != OC) return bail(fail(TB200_ERR_INVALID, "layer %d: concat channels %d != %d", li, coff, OC));
            }
            else if (L.op == TB200_OP_UPSAMPLE)
            {
                s.n = N, s.h = H, s.w_ = W, s.cp_in = tin.cp, s.scale = L.up_scale;
                if (OH != H * L.up_scale || OW != W * L.up_scale) return bail(fail(TB200_ERR_INVALID, "layer %d: upsample shape", li));
            }
            else if (L.op == TB200_OP_IDENTITY || L.op == TB200_OP_CONCAT)
            {
                s.bytes = (long long)(tin.nhwc_bytes / K);
                if (tout.nhwc_bytes != tin.nhwc_bytes) return bail(fail(TB200_ERR_UNSUPPORTED, "layer %d: identity changes the NHWC footprint", li));
            }
            g->layer_kernel[li] = (s.kind == K_CONV_DW && s.dwp.valid) ? "conv_dw3x3_tma_dp4a" : kStepName[s.kind];
            steps.push_back(s);
        }
        for (size_t i = 0; i < g->output_ids.size(); i++)
        {
            TensorInfo& t = g->tensors[g->output_ids[i]];
            Step s;
            s.kind = K_NHWC2NCHW, s.layer = -1, s.in = tdev(t), s.out = g->out_nchw_dev[i] + ck * (t.nchw_bytes / K);
            s.n = nb, s.c = t.d.dims[1], s.h = t.d.dims[2], s.w_ = t.d.dims[3];
            steps.push_back(s);
        }
        return 0;
    };
    // the whole-batch plan serves tb200_graph_launch / profile (device-resident use); the chunk plans serve the pipelined
    // tb200_graph_run.  They address the same tensors (a chunk is a slice of dim 0), so either may run at any time.
    {
        int rc = build_steps(1, 0, g->steps, true);
        if (rc) return rc;
        for (int ck = 0; ck < (int)g->chunk_steps.size(); ck++)
            if ((rc = build_steps(Kpipe, ck, g->chunk_steps[ck], false)) != 0) return rc;
    }
    g->num_launches = (int)g->steps.size();
    CUDA_OK(cudaStreamSynchronize(ctx->stream));

    // ---- capture the launch sequences into CUDA graphs: [0] = whole batch, [1..K] = the pipeline chunks ----
    if (!(flags & TB200_PRERUN_NO_GRAPH))
    {
        const int ngraphs = 1 + (int)g->chunk_steps.size();
        g->cu_graphs.assign(ngraphs, nullptr);
        g->cu_execs.assign(ngraphs, nullptr);
        for (int gi = 0; gi < ngraphs; gi++)
        {
            const std::vector<Step>& seq = gi == 0 ? g->steps : g->chunk_steps[gi - 1];
            CUDA_OK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            int rc = 0;
            for (const Step& s : seq)
                if ((rc = run_step(g, s, ctx->stream)) != 0) break;
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g->cu_graphs[gi]);
            if (rc) return bail(rc);
            if (ce != cudaSuccess) return bail(fail(TB200_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce)));
            ce = cudaGraphInstantiate(&g->cu_execs[gi], g->cu_graphs[gi], 0);
            if (ce != cudaSuccess) return bail(fail(TB200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce)));
        }
        const int K = Kpipe;
        if (K > 1)
        {
            CUDA_OK(cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking));
            CUDA_OK(cudaStreamCreateWithFlags(&g->d2h_stream, cudaStreamNonBlocking));
            g->ev_in.resize(K), g->ev_out.resize(K);
            for (int ck = 0; ck < K; ck++)
            {
                CUDA_OK(cudaEventCreateWithFlags(&g->ev_in[ck], cudaEventDisableTiming));
                CUDA_OK(cudaEventCreateWithFlags(&g->ev_out[ck], cudaEventDisableTiming));
            }
            CUDA_OK(cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming));
        }
    }
    *out = g;
    return 0;
}

extern "C" {

int tb200_graph_upload(tb200_graph* g, int input_index, const void* host_nchw)
{
    if (!g || input_index < 0 || input_index >= (int)g->input_ids.size() || !host_nchw) return fail(TB200_ERR_INVALID, "bad upload arguments");
    CUDA_OK(cudaSetDevice(g->ctx->device));
    CUDA_OK(cudaMemcpyAsync(g->in_nchw_dev[input_index], host_nchw, g->tensors[g->input_ids[input_index]].nchw_bytes,
                            cudaMemcpyHostToDevice, g->ctx->stream));
    return 0;
}

int tb200_graph_launch(tb200_graph* g)
{
    if (!g) return fail(TB200_ERR_INVALID, "null graph");
    CUDA_OK(cudaSetDevice(g->ctx->device));
    if (!g->cu_execs.empty())
    {
        CUDA_OK(cudaGraphLaunch(g->cu_execs[0], g->ctx->stream)); // the whole-batch graph
        return 0;
    }
    for (const Step& s : g->steps)
    {
        int rc = run_step(g, s, g->ctx->stream);
        if (rc) return rc;
    }
    return 0;
}

int tb200_graph_download(tb200_graph* g, int output_index, void* host_nchw)
{
    if (!g || output_index < 0 || output_index >= (int)g->output_ids.size() || !host_nchw) return fail(TB200_ERR_INVALID, "bad download arguments");
    CUDA_OK(cudaSetDevice(g->ctx->device));
    CUDA_OK(cudaMemcpyAsync(host_nchw, g->out_nchw_dev[output_index], g->tensors[g->output_ids[output_index]].nchw_bytes,
                            cudaMemcpyDeviceToHost, g->ctx->stream));
    return 0;
}

int tb200_graph_sync(tb200_graph* g)
{
    if (!g) return fail(TB200_ERR_INVALID, "null graph");
    CUDA_OK(cudaStreamSynchronize(g->ctx->stream));
    return 0;
}

int tb200_graph_run(tb200_graph* g, const void* const* host_inputs, void* const* host_outputs)
{
    if (!g || !host_inputs || !host_outputs) return fail(TB200_ERR_INVALID, "bad run arguments");
    int rc;
    if (!host_inputs || !host_outputs) return fail(TB200_ERR_INVALID, "run: null buffer table");
    for (size_t i = 0; i < g->input_ids.size(); i++)
        if (!host_inputs[i]) return fail(TB200_ERR_INVALID, "run: input %d has no host buffer", (int)i);
    for (size_t i = 0; i < g->output_ids.size(); i++)
        if (!host_outputs[i]) return fail(TB200_ERR_INVALID, "run: output %d has no host buffer (the graph has %d outputs)", (int)i, (int)g->output_ids.size());
    if (g->chunks <= 1 || g->cu_execs.empty())
    {
        for (size_t i = 0; i < g->input_ids.size(); i++)
            if ((rc = tb200_graph_upload(g, (int)i, host_inputs[i])) != 0) return rc;
        if ((rc = tb200_graph_launch(g)) != 0) return rc;
        for (size_t i = 0; i < g->output_ids.size(); i++)
            if ((rc = tb200_graph_download(g, (int)i, host_outputs[i])) != 0) return rc;
        return tb200_graph_sync(g);
    }
    // Pipelined: the H2D copies of all chunks are queued back to back on the copy stream (the link stays busy), chunk k's
    // kernels start as soon as ITS slice has landed, and its outputs leave on a third stream while chunk k+1 computes.
    CUDA_OK(cudaSetDevice(g->ctx->device));
    const int K = g->chunks;
    cudaStream_t cs = g->ctx->stream;
    CUDA_OK(cudaEventRecord(g->ev_done, cs)); // order after whatever the caller queued on the context stream
    CUDA_OK(cudaStreamWaitEvent(g->copy_stream, g->ev_done, 0));
    for (int ck = 0; ck < K; ck++)
    {
        for (size_t i = 0; i < g->input_ids.size(); i++)
        {
            const size_t bytes = g->tensors[g->input_ids[i]].nchw_bytes / K;
            CUDA_OK(cudaMemcpyAsync(g->in_nchw_dev[i] + ck * bytes, (const uint8_t*)host_inputs[i] + ck * bytes, bytes, cudaMemcpyHostToDevice,
                                    g->copy_stream));
        }
        CUDA_OK(cudaEventRecord(g->ev_in[ck], g->copy_stream));
    }
    for (int ck = 0; ck < K; ck++)
    {
        CUDA_OK(cudaStreamWaitEvent(cs, g->ev_in[ck], 0));
        CUDA_OK(cudaGraphLaunch(g->cu_execs[1 + ck], cs));
        CUDA_OK(cudaEventRecord(g->ev_out[ck], cs));
        CUDA_OK(cudaStreamWaitEvent(g->d2h_stream, g->ev_out[ck], 0));
        for (size_t i = 0; i < g->output_ids.size(); i++)
        {
            const size_t bytes = g->tensors[g->output_ids[i]].nchw_bytes / K;
            CUDA_OK(cudaMemcpyAsync((uint8_t*)host_outputs[i] + ck * bytes, g->out_nchw_dev[i] + ck * bytes, bytes, cudaMemcpyDeviceToHost,
                                    g->d2h_stream));
        }
    }
    CUDA_OK(cudaStreamSynchronize(g->d2h_stream));
    CUDA_OK(cudaStreamSynchronize(cs));
    return 0;
}

int tb200_graph_postrun(tb200_graph* g)
{
    if (!g) return 0;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    destroy_graph(g);
    return 0;
}

int tb200_graph_weight_arena(tb200_graph* g, void** device_ptr, size_t* bytes)
{
    if (!g || !device_ptr || !bytes) return fail(TB200_ERR_INVALID, "bad arguments");
    *device_ptr = g->w_arena, *bytes = g->w_bytes;
    return 0;
}

int tb200_graph_num_launches(tb200_graph* g) { return g ? g->num_launches : 0; }

const char* tb200_graph_layer_kernel(tb200_graph* g, int layer)
{
    if (!g || layer < 0 || layer >= (int)g->layer_kernel.size()) return "";
    return g->layer_kernel[layer];
}

int tb200_graph_read_tensor(tb200_graph* g, int tensor_id, void* host_nchw)
{
    if (!g || tensor_id < 0 || tensor_id >= (int)g->tensors.size() || !host_nchw) return fail(TB200_ERR_INVALID, "bad arguments");
    CUDA_OK(cudaSetDevice(g->ctx->device));
    const TensorInfo& t = g->tensors[tensor_id];
    uint8_t* tmp = nullptr;
    CUDA_OK(cudaMalloc(&tmp, t.nchw_bytes));
    const uint8_t* src = t.dev;
    cudaError_t e = launch_nhwc_to_nchw(src, tmp, t.d.dims[0], t.d.dims[1], t.d.dims[2], t.d.dims[3], g->ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_nchw, tmp, t.nchw_bytes, cudaMemcpyDeviceToHost, g->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(TB200_ERR_CUDA, "read_tensor: %s", cudaGetErrorString(e));
    return 0;
}

int tb200_graph_profile(tb200_graph* g, float* layer_ms, int num_layers)
{
    if (!g || !layer_ms || num_layers < (int)g->layers.size()) return fail(TB200_ERR_INVALID, "bad arguments");
    CUDA_OK(cudaSetDevice(g->ctx->device));
    for (int i = 0; i < num_layers; i++) layer_ms[i] = 0.f;
    cudaEvent_t a, b;
    CUDA_OK(cudaEventCreate(&a));
    CUDA_OK(cudaEventCreate(&b));
    for (const Step& s : g->steps)
    {
        cudaEventRecord(a, g->ctx->stream);
        int rc = run_step(g, s, g->ctx->stream);
        cudaEventRecord(b, g->ctx->stream);
        if (rc) return rc;
        CUDA_OK(cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (s.layer >= 0) layer_ms[s.layer] += ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return 0;
}

int tb200_graph_work(tb200_graph* g, double* ops, double* bytes)
{
    if (!g) return fail(TB200_ERR_INVALID, "null graph");
    if (ops) *ops = g->work_ops;
    if (bytes) *bytes = g->work_bytes;
    return 0;
}

// ---- kernel-level entry points (device pointers) -------------------------------------------------------------
static EpiParams epi_from_abi(const tb200k_epilogue* e)
{
    EpiParams p{};
    p.bias = e->bias, p.w_scale = e->w_scale, p.in_scale = e->in_scale, p.out_scale = e->out_scale;
    p.in_zero = e->in_zero, p.w_zero = e->w_zero, p.out_zero = e->out_zero, p.activation = e->activation, p.recipe = e->recipe;
    p.is_uint8 = e->is_uint8, p.fc_rounding = e->fc_rounding, p.has_bias = e->bias != nullptr;
    p.in_w_scale = e->in_scale * e->w_scale_tensor;
    p.fast_ok = 0; // the raw kernel entry points always use the literal reference arithmetic
    return p;
}
static ConvShape shape_from_abi(const tb200k_conv_shape* s)
{
    ConvShape c{};
    c.n = s->n, c.h = s->h, c.w = s->w, c.c = s->c, c.cp = cpad(s->c), c.oh = s->oh, c.ow = s->ow, c.oc = s->oc, c.ocp = cpad(s->oc);
    c.kh = s->kh, c.kw = s->kw, c.sh = s->sh, c.sw = s->sw, c.ph0 = s->ph0, c.pw0 = s->pw0, c.dh = s->dh, c.dw = s->dw, c.group = s->group;
    c.cg = s->c / s->group, c.cgp = s->group == 1 ? c.cp : c.cg;
    return c;
}
#define K_CHECK(cond) \
    if (!(cond)) return fail(TB200_ERR_INVALID, "invalid kernel arguments: %s", #cond)
#define K_LAUNCH(expr)                                                                      \
    do                                                                                      \
    {                                                                                       \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) return fail(TB200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
        return 0;                                                                           \
    } while (0)

int tb200k_conv_direct(const void* in, const void* weight, void* out, const tb200k_conv_shape* s, const tb200k_epilogue* e, void* stream)
{
    K_CHECK(in && weight && out && s && e && s->group >= 1);
    EpiParams p = epi_from_abi(e);
    K_LAUNCH(launch_conv_direct(in, weight, out, shape_from_abi(s), p, (cudaStream_t)stream));
}
int tb200k_conv_dw3x3(const void* in, const void* weight, void* out, const tb200k_conv_shape* s, const tb200k_epilogue* e, void* stream)
{
    K_CHECK(in && weight && out && s && e && s->group == s->c && s->oc == s->c);
    EpiParams p = epi_from_abi(e);
    K_LAUNCH(launch_conv_dw(in, weight, out, shape_from_abi(s), p, (cudaStream_t)stream));
}
int tb200k_conv_stem_nchw(const void* in_nchw, const void* weight, void* out, const tb200k_conv_shape* s, const tb200k_epilogue* e, void* stream)
{
    K_CHECK(in_nchw && weight && out && s && e && s->c <= 4 && s->group == 1);
    EpiParams p = epi_from_abi(e);
    K_LAUNCH(launch_conv_stem(in_nchw, weight, out, shape_from_abi(s), p, (cudaStream_t)stream));
}
int tb200k_gemm_i8(const void* in, const void* weight, void* out, int64_t m, int32_t k_pad, int32_t oc, const tb200k_epilogue* e, void* stream)
{
    K_CHECK(in && weight && out && e && m > 0 && k_pad > 0 && oc > 0 && !e->is_uint8);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fail(TB200_ERR_NO_DEVICE, "no CUDA device");
    GemmPlan plan;
    int rc = gemm_plan_create(&plan, in, k_pad, weight, out, m, k_pad, oc, cpad(oc), cpad(oc), 0, 0);
    if (rc) return fail(rc, "gemm plan failed");
    EpiParams p = epi_from_abi(e);
    K_LAUNCH(launch_gemm_i8(plan, p, nullptr, sms, (cudaStream_t)stream));
}
int tb200k_nchw_to_nhwc(const void* in, void* out, int n, int c, int h, int w, void* stream)
{
    K_CHECK(in && out && n > 0 && c > 0 && h > 0 && w > 0);
    K_LAUNCH(launch_nchw_to_nhwc(in, out, n, c, h, w, (cudaStream_t)stream));
}
int tb200k_nhwc_to_nchw(const void* in, void* out, int n, int c, int h, int w, void* stream)
{
    K_CHECK(in && out && n > 0 && c > 0 && h > 0 && w > 0);
    K_LAUNCH(launch_nhwc_to_nchw(in, out, n, c, h, w, (cudaStream_t)stream));
}

} // extern "C"

// b200_device.cc -- the Tengine nn_device "B200": source/device/b200/ of the integration build.
//
// This is the drop-in boundary of SURVEY.md 8(b): one C-linkage pair register_b200_device() /
// unregister_b200_device() (the names cmake/registry.cmake:12-15 derives from this file's stem) that hands a
// `struct device {name, interface, allocator, optimizer}` (source/device/device.h:41-108) to register_device()
// (source/module/module.c:280).  Everything below the vtables goes through the plain C ABI of
// include/tengine_b200.h; no kernel code lives here and nothing runs on the CPU: a node this device accepts is
// executed by libtengine_b200.so on the GPU or the graph fails (return < 0, TLOG_ERR; c_api.c marks GRAPH_STAT_ERROR).
//
// init_tengine()/create_graph()/run_graph(), the tmfile serializer, tm_classification_int8/uint8 and tm_benchmark
// are untouched.  Apps select the device with set_context_device(ctx, "B200", ...) (tm_benchmark -d B200) or, for
// binaries that pass no context (examples/tm_classification_int8.c:83), with the environment variable
// TG_DEFAULT_DEVICE=B200: registration then replaces the `optimizer` of the already registered CPU device by one
// whose split_graph performs this device's split (CPU keeps its interface, so CPU-side subgraphs still run there).
extern "C" {
#include "api/c_api.h"
#include "device/device.h"
#include "graph/graph.h"
#include "graph/node.h"
#include "graph/subgraph.h"
#include "graph/tensor.h"
#include "executer/executer.h"
#include "module/module.h"
#include "operator/op.h"
#include "optimizer/split.h"
#include "utility/log.h"
#include "utility/sys_port.h"
#include "utility/vector.h"
#include "operator/prototype/convolution_param.h"
#include "operator/prototype/pooling_param.h"
#include "operator/prototype/fc_param.h"
#include "operator/prototype/relu_param.h"
#include "operator/prototype/eltwise_param.h"
#include "operator/prototype/concat_param.h"
#include "operator/prototype/upsample_param.h"
#include "operator/prototype/softmax_param.h"
}

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "tengine_b200.h"

#define B200_DEV_NAME "B200"

namespace {

const int kSupportedOps[] = {OP_CONST,  OP_INPUT,    OP_CONV,    OP_FC,      OP_POOL,    OP_RELU,      OP_ELTWISE, OP_CONCAT,
                             OP_UPSAMPLE, OP_DROPOUT, OP_SOFTMAX, OP_SIGMOID, OP_FLATTEN, OP_RESHAPE, OP_HARDSWISH};

struct B200Graph
{
    tb200_graph* graph = nullptr;
    std::vector<uint16_t> inputs, outputs; // ir tensor ids in ABI order
};

struct B200Device
{
    struct device base;
    tb200_context* ctx;
};

tb200_context* g_ctx = nullptr;
std::vector<int> g_ctx_devs; // the CUDA ordinals g_ctx was created over
int g_live_graphs = 0;       // device graphs alive on g_ctx

int b200_dev_init(struct device* dev)
{
    (void)dev;
    return 0; // the GPU is bound lazily at the first pre_run so that init_tengine() works on GPU-less hosts
}

// The GPUs behind the device: the option blob of set_context_device(ctx, "B200", &opt, sizeof opt) (tb200_device_option, the
// analogue of trt_option, source/device/tensorrt/trt_define.h:36-42) or, for applications that pass none (tm_benchmark,
// tm_classification_*), the environment: TG_B200_GPUS=<count> (or "all"), TG_B200_GPU=<first ordinal>.  One process, one
// application thread; the batch of every run is sharded over the group inside libtengine_b200.so.
int ensure_context(const void* options)
{
    int first = 0, count = 1;
    if (const char* e = getenv("TG_B200_GPU")) first = atoi(e);
    if (const char* e = getenv("TG_B200_GPUS")) count = (0 == strcmp(e, "all")) ? tb200_device_count() : atoi(e);
    if (options)
    {
        const tb200_device_option* o = (const tb200_device_option*)options;
        if (o->dev_name && 0 == strcmp(o->dev_name, B200_DEV_NAME))
        {
            if (o->num_gpus > 0) count = o->num_gpus;
            if (o->first_gpu > 0) first = o->first_gpu;
        }
    }
    if (count < 1) count = 1;
    std::vector<int> devs;
    for (int i = 0; i < count; i++) devs.push_back(first + i);
    if (const char* e = getenv("TG_B200_GPU_LIST")) // explicit ordinals, e.g. "0,2,4,6"; a repeated ordinal = several shards on one GPU (tests)
    {
        devs.clear();
        for (const char* p = e; *p;)
        {
            char* end;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            devs.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
        if (devs.empty()) devs.push_back(first);
        count = (int)devs.size(), first = devs[0];
    }
    if (g_ctx && devs == g_ctx_devs) return 0;
    if (g_ctx)
    {
        // another graph asks for a different group (e.g. a second context with another option blob)
        if (g_live_graphs > 0)
        {
            TLOG_ERR("Tengine: B200 device: a graph is still prepared on another GPU group; postrun it first\n");
            return -1;
        }
        tb200_context_destroy(g_ctx);
        g_ctx = nullptr;
    }
    g_ctx_devs = devs;
    const int rc = count == 1 ? tb200_context_create(first, &g_ctx) : tb200_context_create_multi(devs.data(), count, &g_ctx);
    if (rc != 0)
    {
        TLOG_ERR("Tengine: B200 device: %s\n", tb200_last_error());
        return -1;
    }
    return 0;
}

// ---- what the engine can take, node by node (the planner's own predicates, tengine_b200/csrc/engine.cu).  An op type with a
//      node that fails its predicate is left to the CPU device for this graph (the splitter works on op types, split.c:140). ----
bool node_supported(struct graph* ir_graph, struct node* node)
{
    struct tensor* in0 = node->input_num ? get_ir_graph_tensor(ir_graph, node->input_tensors[0]) : nullptr;
    struct tensor* out0 = node->output_num ? get_ir_graph_tensor(ir_graph, node->output_tensors[0]) : nullptr;
    switch (node->op.type)
    {
    case OP_CONV:
    {
        const struct conv_param* p = (const struct conv_param*)node->op.param_mem;
        if (!in0 || in0->dim_num != 4) return false;
        const int cg = in0->dims[1] / (p->group > 0 ? p->group : 1), og = out0->dims[1] / (p->group > 0 ? p->group : 1);
        const bool depthwise = p->group > 1 && cg == 1 && og == 1;
        if (p->group > 1 && !depthwise && ((cg % 4) || (og % 4))) return false;
        return true;
    }
    case OP_FC:
    {
        struct tensor* w = get_ir_graph_tensor(ir_graph, node->input_tensors[1]);
        return w->dim_num == 2 && out0 && w->dims[0] == out0->dims[1]; // the [K, N] layout (fc_ref.c need_trans) stays on the CPU
    }
    case OP_POOL: return in0 && in0->dim_num == 4;
    case OP_ELTWISE:
    {
        const int t = ((const struct eltwise_param*)node->op.param_mem)->type;
        if (node->input_num != 2 || (t != ELT_SUM && t != ELT_PROD)) return false;
        struct tensor* in1 = get_ir_graph_tensor(ir_graph, node->input_tensors[1]);
        return in1->tensor_type != TENSOR_TYPE_CONST && in1->elem_num == in0->elem_num;
    }
    case OP_CONCAT: return in0 && in0->dim_num == 4 && node->input_num <= 4 && ((const struct concat_param*)node->op.param_mem)->axis == 1;
    case OP_UPSAMPLE:
    {
        const float sc = ((const struct upsample_param*)node->op.param_mem)->scale;
        return in0 && in0->dim_num == 4 && sc == (float)(int)sc && sc >= 1.f;
    }
    case OP_SOFTMAX:
    {
        int axis = ((const struct softmax_param*)node->op.param_mem)->axis;
        if (in0 && axis < 0) axis += in0->dim_num;
        return axis == 1;
    }
    case OP_HARDSWISH: return in0 && in0->data_type == TENGINE_DT_UINT8; // the reference has no int8 kernel (hardswish_ref.c:59-66)
    case OP_FLATTEN:
    case OP_RESHAPE: return in0 && out0 && in0->dims[0] == out0->dims[0] && in0->dim_num <= 4 && out0->dim_num <= 4;
    default: return true;
    }
}

void fill_tensor_desc(const struct tensor* t, tb200_tensor_desc* d)
{
    d->data_type = t->data_type;
    for (int k = 0; k < 4; k++) d->dims[k] = k < t->dim_num ? t->dims[k] : 1;
    d->scale = t->scale; // per-tensor (quant_param_num == 1): the union holds the scalar (tensor.h:78-92)
    d->zero_point = t->zero_point;
}

// Which float recipe the reference CPU device would have used for this convolution
// (score-based selection, source/device/cpu/cpu_module.c:135-170).
int conv_recipe(const struct conv_param* p, const struct tensor* in, const struct tensor* out)
{
    const bool depthwise = p->group > 1 && p->group == in->dims[1] && p->group == out->dims[1];
    if (p->group == 1) return TB200_RECIPE_HCL; // conv_hcl_x86 / conv_direct_hcl_int8_x86
    if (depthwise && in->data_type == TENGINE_DT_INT8 && in->dims[0] == 1 && p->kernel_h == 3 && p->kernel_w == 3 &&
        p->stride_h == p->stride_w && (p->stride_h == 1 || p->stride_h == 2) && p->dilation_h == 1 && p->dilation_w == 1 &&
        p->pad_h0 == p->pad_h1 && p->pad_w0 == p->pad_w1)
        return TB200_RECIPE_HCL; // conv_dw_hcl_x86.c:508-543
    return TB200_RECIPE_REF;     // conv_ref.c
}

int b200_dev_prerun(struct device* dev, struct subgraph* subgraph, void* options)
{
    (void)dev;
    struct graph* ir_graph = subgraph->graph;

    std::map<uint16_t, int> tmap; // ir tensor id -> index in the ABI tensor table
    std::vector<tb200_tensor_desc> tensors;
    std::vector<tb200_layer_desc> layers;
    std::vector<std::vector<float>> scale_store; // expanded per-channel scales, kept alive until prerun returns
    auto tensor_id = [&](uint16_t ir_id) {
        auto it = tmap.find(ir_id);
        if (it != tmap.end()) return it->second;
        tb200_tensor_desc d;
        fill_tensor_desc(get_ir_graph_tensor(ir_graph, ir_id), &d);
        tensors.push_back(d);
        tmap[ir_id] = (int)tensors.size() - 1;
        return (int)tensors.size() - 1;
    };

    for (int i = 0; i < subgraph->node_num; i++)
    {
        struct node* node = get_ir_graph_node(ir_graph, subgraph->node_list[i]);
        const int op = node->op.type;
        if (op == OP_CONST || op == OP_INPUT) continue;
        tb200_layer_desc L;
        memset(&L, 0, sizeof L);
        L.kernel_h = L.kernel_w = L.stride_h = L.stride_w = L.dilation_h = L.dilation_w = L.group = 1;
        L.activation = -1, L.axis = 1, L.up_scale = 1, L.elt_type = TB200_ELT_SUM;
        struct tensor* in0 = get_ir_graph_tensor(ir_graph, node->input_tensors[0]);
        struct tensor* out0 = get_ir_graph_tensor(ir_graph, node->output_tensors[0]);
        L.output = tensor_id(node->output_tensors[0]);
        L.num_inputs = 1;
        L.inputs[0] = tensor_id(node->input_tensors[0]);
        switch (op)
        {
        case OP_CONV:
        case OP_FC:
        {
            struct tensor* w = get_ir_graph_tensor(ir_graph, node->input_tensors[1]);
            struct tensor* b = node->input_num > 2 ? get_ir_graph_tensor(ir_graph, node->input_tensors[2]) : nullptr;
            const int oc = out0->dims[1];
            L.weight = w->data;
            L.bias = b ? (const int32_t*)b->data : nullptr;
            L.bias_scale = b ? (b->quant_param_num > 1 ? b->scale_list[0] : b->scale) : 0.f;
            if (in0->data_type == TENGINE_DT_INT8)
            {
                scale_store.emplace_back(oc);
                std::vector<float>& s = scale_store.back();
                for (int c = 0; c < oc; c++) s[c] = w->quant_param_num > 1 ? w->scale_list[c] : w->scale;
                L.weight_scales = s.data();
                L.weight_zero = 0;
            }
            else
            {
                scale_store.emplace_back(1, w->quant_param_num > 1 ? w->scale_list[0] : w->scale);
                L.weight_scales = scale_store.back().data();
                L.weight_zero = w->quant_param_num > 1 ? w->zp_list[0] : w->zero_point;
            }
            if (op == OP_CONV)
            {
                const struct conv_param* p = (const struct conv_param*)node->op.param_mem;
                L.op = TB200_OP_CONV;
                L.kernel_h = p->kernel_h, L.kernel_w = p->kernel_w, L.stride_h = p->stride_h, L.stride_w = p->stride_w;
                L.pad_h0 = p->pad_h0, L.pad_h1 = p->pad_h1, L.pad_w0 = p->pad_w0, L.pad_w1 = p->pad_w1;
                L.dilation_h = p->dilation_h, L.dilation_w = p->dilation_w, L.group = p->group, L.activation = p->activation;
                L.recipe = conv_recipe(p, in0, out0);
            }
            else
            {
                L.op = TB200_OP_FC;
            }
            break;
        }
        case OP_POOL:
        {
            const struct pool_param* p = (const struct pool_param*)node->op.param_mem;
            L.op = TB200_OP_POOL;
            L.pool_method = p->pool_method, L.pool_global = p->global, L.caffe_flavor = p->caffe_flavor;
            L.kernel_h = p->kernel_h, L.kernel_w = p->kernel_w, L.stride_h = p->stride_h, L.stride_w = p->stride_w;
            L.pad_h0 = p->pad_h0, L.pad_h1 = p->pad_h1, L.pad_w0 = p->pad_w0, L.pad_w1 = p->pad_w1;
            break;
        }
        case OP_RELU:
            L.op = TB200_OP_RELU;
            L.negative_slope = ((const struct relu_param*)node->op.param_mem)->negative_slope;
            break;
        case OP_ELTWISE:
            L.op = TB200_OP_ELTWISE;
            L.elt_type = ((const struct eltwise_param*)node->op.param_mem)->type;
            L.num_inputs = node->input_num;
            if (node->input_num != 2)
            {
                TLOG_ERR("Tengine: B200 device: eltwise with %d inputs is not supported\n", node->input_num);
                return -1;
            }
            L.inputs[1] = tensor_id(node->input_tensors[1]);
            break;
        case OP_CONCAT:
            L.op = TB200_OP_CONCAT;
            L.axis = ((const struct concat_param*)node->op.param_mem)->axis;
            if (node->input_num > 4)
            {
                TLOG_ERR("Tengine: B200 device: concat with %d inputs is not supported\n", node->input_num);
                return -1;
            }
            L.num_inputs = node->input_num;
            for (int k = 1; k < node->input_num; k++) L.inputs[k] = tensor_id(node->input_tensors[k]);
            break;
        case OP_UPSAMPLE:
            L.op = TB200_OP_UPSAMPLE;
            L.up_scale = (int)((const struct upsample_param*)node->op.param_mem)->scale;
            break;
        case OP_DROPOUT: L.op = TB200_OP_IDENTITY; break;
        case OP_SOFTMAX:
            L.op = TB200_OP_SOFTMAX;
            L.axis = ((const struct softmax_param*)node->op.param_mem)->axis;
            if (L.axis < 0) L.axis += in0->dim_num;
            break;
        case OP_SIGMOID: L.op = TB200_OP_SIGMOID; break;
        case OP_HARDSWISH: L.op = TB200_OP_HARDSWISH; break;
        case OP_FLATTEN:
        case OP_RESHAPE: L.op = TB200_OP_RESHAPE; break; // the output tensor's dims carry the new shape
        default: TLOG_ERR("Tengine: B200 device: op %d reached pre_run but is not supported\n", op); return -1;
        }
        layers.push_back(L);
    }

    B200Graph* bg = new B200Graph();
    std::vector<int32_t> in_ids, out_ids;
    for (int i = 0; i < subgraph->input_num; i++)
    {
        struct tensor* t = get_ir_graph_tensor(ir_graph, subgraph->input_tensor_list[i]);
        if (t->tensor_type != TENSOR_TYPE_VAR && t->tensor_type != TENSOR_TYPE_INPUT) continue;
        bg->inputs.push_back(subgraph->input_tensor_list[i]);
        in_ids.push_back(tensor_id(subgraph->input_tensor_list[i]));
    }
    for (int i = 0; i < subgraph->output_num; i++)
    {
        bg->outputs.push_back(subgraph->output_tensor_list[i]);
        out_ids.push_back(tensor_id(subgraph->output_tensor_list[i]));
    }
    if (const char* dump = getenv("TG_B200_DUMP_DESC"))
    {
        // debug: the descriptor tables exactly as they go to tb200_graph_prerun (text; constants as FNV-1a hashes)
        if (FILE* f = fopen(dump, "a"))
        {
            fprintf(f, "# subgraph %d: %d nodes, %d inputs, %d outputs\n", subgraph->index, subgraph->node_num, subgraph->input_num, subgraph->output_num);
            auto fnv = [](const void* p, size_t n) {
                uint64_t h = 1469598103934665603ull;
                for (size_t i = 0; i < n; i++) h = (h ^ ((const uint8_t*)p)[i]) * 1099511628211ull;
                return (unsigned long long)h;
            };
            for (size_t i = 0; i < tensors.size(); i++)
                fprintf(f, "T%zu dt%d [%d %d %d %d] scale %.9g zp %d\n", i, tensors[i].data_type, tensors[i].dims[0], tensors[i].dims[1], tensors[i].dims[2],
                        tensors[i].dims[3], tensors[i].scale, tensors[i].zero_point);
            for (size_t i = 0; i < layers.size(); i++)
            {
                const tb200_layer_desc& L = layers[i];
                const tb200_tensor_desc& to = tensors[L.output];
                const tb200_tensor_desc& ti = tensors[L.inputs[0]];
                size_t wn = 0;
                if (L.op == TB200_OP_CONV) wn = (size_t)to.dims[1] * (ti.dims[1] / L.group) * L.kernel_h * L.kernel_w;
                if (L.op == TB200_OP_FC) wn = (size_t)to.dims[1] * ti.dims[1] * ti.dims[2] * ti.dims[3];
                fprintf(f, "L%zu op%d in[%d %d %d %d]/%d out%d k%dx%d s%d,%d p%d,%d,%d,%d d%d,%d g%d act%d rec%d pool%d,%d,%d slope%.9g elt%d axis%d up%d wz%d bs%.9g w%llx b%llx ws%llx\n",
                        i, L.op, L.inputs[0], L.inputs[1], L.inputs[2], L.inputs[3], L.num_inputs, L.output, L.kernel_h, L.kernel_w, L.stride_h, L.stride_w, L.pad_h0,
                        L.pad_h1, L.pad_w0, L.pad_w1, L.dilation_h, L.dilation_w, L.group, L.activation, L.recipe, L.pool_method, L.pool_global, L.caffe_flavor,
                        L.negative_slope, L.elt_type, L.axis, L.up_scale, L.weight_zero, L.bias_scale, wn ? fnv(L.weight, wn) : 0ull,
                        (wn && L.bias) ? fnv(L.bias, (size_t)to.dims[1] * 4) : 0ull,
                        wn ? fnv(L.weight_scales, ti.data_type == TB200_DT_UINT8 ? 4 : (size_t)to.dims[1] * 4) : 0ull);
            }
            fclose(f);
        }
    }
    if (ensure_context(options) != 0) // options: NULL for apps that never call set_context_device (scheduler.c:49-57)
    {
        delete bg;
        return -1;
    }
    int flags = TB200_PRERUN_DEFAULT;
    if (getenv("TG_B200_NO_TENSORCORE")) flags |= TB200_PRERUN_NO_TENSORCORE;
    int rc = tb200_graph_prerun(g_ctx, tensors.data(), (int)tensors.size(), layers.data(), (int)layers.size(), in_ids.data(),
                                (int)in_ids.size(), out_ids.data(), (int)out_ids.size(), flags, &bg->graph);
    if (rc != 0)
    {
        TLOG_ERR("Tengine: B200 device: pre_run of subgraph %d failed: %s\n", subgraph->index, tb200_last_error());
        delete bg;
        return -1;
    }
    subgraph->device_graph = bg;
    g_live_graphs++;
    return 0;
}

int b200_dev_run(struct device* dev, struct subgraph* subgraph)
{
    (void)dev;
    B200Graph* bg = (B200Graph*)subgraph->device_graph;
    if (!bg) return -1;
    struct graph* ir_graph = subgraph->graph;
    std::vector<const void*> ins;
    std::vector<void*> outs;
    for (uint16_t id : bg->inputs)
    {
        struct tensor* t = get_ir_graph_tensor(ir_graph, id); // the application's buffer: re-read every run (c_api.c:1141-1160)
        if (!t->data)
        {
            TLOG_ERR("Tengine: B200 device: input tensor %s has no buffer\n", t->name);
            return -1;
        }
        ins.push_back(t->data);
    }
    for (uint16_t id : bg->outputs)
    {
        struct tensor* t = get_ir_graph_tensor(ir_graph, id);
        if (!t->data) // as source/device/tim-vx/timvx_executor.cc:638-644
        {
            t->data = sys_malloc((size_t)t->elem_num * t->elem_size);
            t->free_host_mem = 1;
            t->internal_allocated = 0;
        }
        outs.push_back(t->data);
    }
    if (tb200_graph_run(bg->graph, ins.data(), outs.data()) != 0)
    {
        TLOG_ERR("Tengine: B200 device: run failed: %s\n", tb200_last_error());
        return -1;
    }
    return 0;
}

int b200_dev_postrun(struct device* dev, struct subgraph* subgraph)
{
    (void)dev;
    B200Graph* bg = (B200Graph*)subgraph->device_graph;
    if (bg)
    {
        tb200_graph_postrun(bg->graph);
        g_live_graphs--;
        delete bg;
        subgraph->device_graph = nullptr;
    }
    return 0;
}

int b200_dev_release(struct device* dev)
{
    (void)dev;
    if (g_ctx) tb200_context_destroy(g_ctx), g_ctx = nullptr;
    return 0;
}

int b200_describe(struct device* device, struct vector* allowed_ops, struct vector* blocked_ops, struct vector* precision)
{
    (void)device;
    for (int op : kSupportedOps) push_vector_data(allowed_ops, &op);
    for (int i = 0; i < OP_BUILTIN_LAST; i++)
    {
        bool in_list = false;
        for (int op : kSupportedOps) in_list |= (op == i);
        if (!in_list) push_vector_data(blocked_ops, &i);
    }
    int p = TENGINE_DT_INT8;
    push_vector_data(precision, &p);
    p = TENGINE_DT_UINT8;
    push_vector_data(precision, &p);
    return 0;
}

int b200_evaluation(struct device*, struct subgraph*, struct vector*, struct vector*) { return 0; }

int b200_allocate(struct device* device, struct subgraph* sub_graph)
{
    if (!device) return -1;
    // the scheduler waits for input_wait_count producers (scheduler.c:117-125): INPUT tensors are always ready
    sub_graph->input_wait_count = 0;
    for (int i = 0; i < sub_graph->input_num; i++)
    {
        struct tensor* t = get_ir_graph_tensor(sub_graph->graph, sub_graph->input_tensor_list[i]);
        if (t->tensor_type == TENSOR_TYPE_VAR) sub_graph->input_wait_count++;
    }
    return 0;
}

int b200_release(struct device*, struct subgraph*) { return 0; }

extern struct B200Device g_b200;

int b200_split_graph(struct graph* ir_graph)
{
    // reached either because the context names this device, or through the TG_DEFAULT_DEVICE seam on the CPU device;
    // the splitter assigns context->device to accelerator subgraphs without a NULL check (split.c:214-215)
    ir_graph->attribute->context->device = &g_b200.base;
    struct vector* allowed_all = create_vector(sizeof(int), nullptr);
    struct vector* blocked_all = create_vector(sizeof(int), nullptr);
    struct vector* precision = create_vector(sizeof(int), nullptr);
    b200_describe(&g_b200.base, allowed_all, blocked_all, precision);
    // op types with a node the engine cannot take go to the CPU device for this graph (instead of failing pre_run)
    std::vector<char> veto(OP_BUILTIN_LAST, 0);
    for (int i = 0; i < ir_graph->node_num; i++)
    {
        struct node* node = get_ir_graph_node(ir_graph, i);
        if (node->op.type < OP_BUILTIN_LAST && !node_supported(ir_graph, node)) veto[node->op.type] = 1;
    }
    struct vector* allowed_ops = create_vector(sizeof(int), nullptr);
    struct vector* blocked_ops = create_vector(sizeof(int), nullptr);
    for (int i = 0; i < OP_BUILTIN_LAST; i++)
    {
        bool ok = false;
        for (int op : kSupportedOps) ok |= (op == i);
        push_vector_data((ok && !veto[i]) ? allowed_ops : blocked_ops, &i);
    }
    release_vector(allowed_all);
    release_vector(blocked_all);
    split_graph_node_to_sub_graph(ir_graph, allowed_ops, blocked_ops, precision);
    release_vector(allowed_ops);
    release_vector(blocked_ops);
    release_vector(precision);
    generate_sub_graph_io(ir_graph);
    add_sub_graph_to_ir_graph(ir_graph);
    for (int i = 0; i < (uint16_t)get_vector_num(ir_graph->subgraph_list); i++)
    {
        struct subgraph* sub_graph = *(struct subgraph**)get_vector_data(ir_graph->subgraph_list, i);
        sub_graph->index = i;
        for (uint16_t j = 0; j < sub_graph->node_num; j++) get_ir_graph_node(ir_graph, sub_graph->node_list[j])->subgraph_idx = sub_graph->index;
    }
    return 0;
}

struct interface b200_interface = {b200_dev_init, b200_dev_prerun, b200_dev_run, b200_dev_postrun, nullptr, nullptr, nullptr, b200_dev_release};
struct allocator b200_allocator = {b200_describe, b200_evaluation, b200_allocate, b200_release};
struct optimizer b200_optimizer = {b200_split_graph, nullptr};
struct B200Device g_b200 = {{B200_DEV_NAME, &b200_interface, &b200_allocator, &b200_optimizer, nullptr, nullptr}, nullptr};

} // namespace

extern "C" {

int register_b200_device(void)
{
    if (register_device(&g_b200.base) != 0)
    {
        TLOG_INFO("Tengine plugin %s register failed.\n", g_b200.base.name);
        return -1;
    }
    // default-device seam: runs after register_cpu_device() because _REGISTER_DEVICE_LIST is ordered
    const char* want = getenv("TG_DEFAULT_DEVICE");
    if (want && 0 == strcmp(want, B200_DEV_NAME))
    {
        struct device* cpu = find_device_via_name("CPU");
        if (cpu) cpu->optimizer = &b200_optimizer;
    }
    return 0;
}

int unregister_b200_device(void) { return unregister_device(&g_b200.base); }
}

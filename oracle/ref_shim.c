#define _GNU_SOURCE
#include <sched.h>
/*
 * ref_shim.c -- drive the UNMODIFIED reference (oracle/_ref/libtengine-lite.so) from this repo's layer
 * descriptors.  TEST INFRASTRUCTURE ONLY (see oracle/tb200_oracle.c header).
 *
 * It builds a Tengine graph through the reference's public C API exactly the way the reference's own
 * per-device op tests do (tests/op/test_op.h:619-664 create_common_test_graph, tests/op/
 * test_timvx_op_convolution.cpp:32-93: create_graph_node + Const nodes + node->op.param_mem), runs it on a
 * named device (default: the reference CPU device) and returns the requested tensors.  It is used to
 *   (1) pin oracle/tb200_oracle.c against the real reference on seeded inputs,
 *   (2) generate tests/golden/ fixtures (tests/golden/make_golden.py),
 *   (3) time the reference CPU backend for bench.py --impl reference / cpu_baseline,
 *   (4) run the same graph on the "B200" device when the integration library is loaded instead.
 * Compiled against the reference headers where they lie (never copied): see oracle/Makefile.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "api/c_api.h"
#include "graph/graph.h"
#include "graph/node.h"
#include "graph/tensor.h"
#include "operator/prototype/convolution_param.h"
#include "operator/prototype/pooling_param.h"
#include "operator/prototype/fc_param.h"
#include "operator/prototype/relu_param.h"
#include "operator/prototype/eltwise_param.h"
#include "operator/prototype/concat_param.h"
#include "operator/prototype/upsample_param.h"
#include "operator/prototype/softmax_param.h"
#include "operator/prototype/flatten_param.h"
#include "operator/op.h"
#include <unistd.h>

#include "../include/tengine_b200.h"

#define SHIM_API __attribute__((visibility("default")))

static int g_inited = 0;

static double now_ms(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1000.0 + tv.tv_usec / 1000.0;
}

static tensor_t make_const(graph_t graph, const char* name, int dtype, const int* dims, int ndim, const void* data,
                           int bytes, const float* scales, const int* zps, int nq)
{
    node_t node = create_graph_node(graph, name, "Const");
    tensor_t t = create_graph_tensor(graph, name, dtype);
    if (!node || !t) return NULL;
    set_node_output_tensor(node, 0, t, TENSOR_TYPE_CONST);
    set_tensor_shape(t, dims, ndim);
    if (set_tensor_buffer(t, (void*)data, bytes) < 0) return NULL;
    if (nq > 0) set_tensor_quant_param(t, scales, zps, nq);
    return t;
}

struct built
{
    graph_t graph;
    context_t ctx;
    tensor_t* tt;
    char** out_node_name;
    int precision;
};

static void free_built(struct built* b, int num_tensors)
{
    if (b->graph) destroy_graph(b->graph);
    if (b->ctx) destroy_context(b->ctx);
    if (b->out_node_name)
        for (int i = 0; i < num_tensors; i++) free(b->out_node_name[i]);
    free(b->out_node_name);
    free(b->tt);
    memset(b, 0, sizeof *b);
}

/* Build the Tengine graph for (tensors, layers).  `want_ids` tensors are marked graph outputs; when out_bufs is given
 * they also receive the caller's buffers before prerun. */
static int build_graph(struct built* B, const tb200_tensor_desc* tensors, int num_tensors, const tb200_layer_desc* layers,
                       int num_layers, const int* input_ids, int num_inputs, const int* want_ids, int num_want,
                       void* const* out_bufs, const char* device_name)
{
    int rc = -1;
    memset(B, 0, sizeof *B);

    if (!g_inited)
    {
        if (init_tengine() != 0) return -100;
        g_inited = 1;
    }
    context_t ctx = NULL;
    if (device_name && device_name[0] && strcmp(device_name, "CPU") != 0)
    {
        ctx = create_context("shim", 1);
        if (set_context_device(ctx, device_name, NULL, 0) < 0)
        {
            fprintf(stderr, "ref_shim: device %s not registered\n", device_name);
            return -101;
        }
    }
    graph_t graph = create_graph(ctx, NULL, NULL);
    if (!graph) return -102;
    set_graph_layout(graph, TENGINE_LAYOUT_NCHW);

    tensor_t* tt = (tensor_t*)calloc(num_tensors, sizeof(tensor_t));
    char** out_node_name = (char**)calloc(num_tensors, sizeof(char*));
    char name[64];
    int precision = TENGINE_MODE_INT8;

    for (int i = 0; i < num_inputs; i++)
    {
        const tb200_tensor_desc* d = &tensors[input_ids[i]];
        snprintf(name, sizeof name, "in%d", input_ids[i]);
        node_t node = create_graph_node(graph, name, "InputOp");
        tensor_t t = create_graph_tensor(graph, name, d->data_type);
        if (!node || !t) goto fail;
        set_node_output_tensor(node, 0, t, TENSOR_TYPE_INPUT);
        set_tensor_shape(t, d->dims, 4);
        set_tensor_quant_param(t, &d->scale, &d->zero_point, 1);
        tt[input_ids[i]] = t;
        out_node_name[input_ids[i]] = strdup(name);
        if (d->data_type == TENGINE_DT_UINT8) precision = TENGINE_MODE_UINT8;
    }

    for (int li = 0; li < num_layers; li++)
    {
        const tb200_layer_desc* L = &layers[li];
        const tb200_tensor_desc* din = &tensors[L->inputs[0]];
        const tb200_tensor_desc* dout = &tensors[L->output];
        const char* opname = NULL;
        switch (L->op)
        {
        case TB200_OP_CONV: opname = "Convolution"; break;
        case TB200_OP_FC: opname = "FullyConnected"; break;
        case TB200_OP_POOL: opname = "Pooling"; break;
        case TB200_OP_RELU: opname = "ReLU"; break;
        case TB200_OP_ELTWISE: opname = "Eltwise"; break;
        case TB200_OP_CONCAT: opname = "Concat"; break;
        case TB200_OP_UPSAMPLE: opname = "Upsample"; break;
        case TB200_OP_IDENTITY: opname = "Dropout"; break;
        case TB200_OP_SOFTMAX: opname = "Softmax"; break;
        case TB200_OP_SIGMOID: opname = "Sigmoid"; break;
        case TB200_OP_HARDSWISH: opname = "Hardswish"; break;
        case TB200_OP_RESHAPE: opname = "Flatten"; break; /* [N,C,H,W] -> [N,C*H*W,1,1]: the only reshape the five graphs contain */
        default: goto fail;
        }
        /* Const nodes are created BEFORE the node that consumes them: the tmfile writer and the graph splitter assume
         * node indices are in topological order (tools/save_graph/save_graph.cpp:233-244, optimizer/split.c:146-160). */
        tensor_t wt = NULL, bt = NULL;
        if (L->op == TB200_OP_CONV || L->op == TB200_OP_FC)
        {
            const int is_u8 = din->data_type == TENGINE_DT_UINT8;
            const int oc = dout->dims[1];
            int wdims[4], wn;
            int wbytes;
            if (L->op == TB200_OP_CONV)
            {
                wdims[0] = oc, wdims[1] = din->dims[1] / L->group, wdims[2] = L->kernel_h, wdims[3] = L->kernel_w;
                wn = 4;
                wbytes = wdims[0] * wdims[1] * wdims[2] * wdims[3];
            }
            else
            {
                wdims[0] = oc, wdims[1] = din->dims[1] * din->dims[2] * din->dims[3];
                wn = 2;
                wbytes = wdims[0] * wdims[1];
            }
            int* zps = (int*)calloc(oc, sizeof(int));
            float* bscales = (float*)calloc(oc, sizeof(float));
            snprintf(name, sizeof name, "L%d_w", li);
            if (is_u8)
            {
                int wz = L->weight_zero;
                wt = make_const(graph, name, TENGINE_DT_UINT8, wdims, wn, L->weight, wbytes, L->weight_scales, &wz, 1);
                bscales[0] = (L->op == TB200_OP_FC && L->bias_scale != 0.f) ? L->bias_scale : din->scale * L->weight_scales[0];
            }
            else
            {
                wt = make_const(graph, name, TENGINE_DT_INT8, wdims, wn, L->weight, wbytes, L->weight_scales, zps, oc);
                for (int c = 0; c < oc; c++) bscales[c] = din->scale * L->weight_scales[c];
            }
            if (!wt) goto fail;
            if (L->bias)
            {
                snprintf(name, sizeof name, "L%d_b", li);
                int bd[1] = {oc};
                bt = make_const(graph, name, TENGINE_DT_INT32, bd, 1, L->bias, oc * 4, bscales, zps, is_u8 ? 1 : oc);
                if (!bt) goto fail;
            }
            free(zps);
            free(bscales);
        }
        snprintf(name, sizeof name, "L%d", li);
        struct node* node = (struct node*)create_graph_node(graph, name, opname);
        if (!node) goto fail;
        for (int k = 0; k < L->num_inputs; k++)
        {
            if (!tt[L->inputs[k]]) { fprintf(stderr, "ref_shim: layer %d input %d undefined\n", li, k); goto fail; }
            set_node_input_tensor(node, k, tt[L->inputs[k]]);
        }
        if (wt) set_node_input_tensor(node, 1, wt);
        if (bt) set_node_input_tensor(node, 2, bt);
        tensor_t ot = create_graph_tensor(graph, name, dout->data_type);
        set_node_output_tensor(node, 0, ot, TENSOR_TYPE_VAR);
        set_tensor_quant_param(ot, &dout->scale, &dout->zero_point, 1);
        /* Requested tensors get the caller's buffer BEFORE prerun: the CPU device's memory pool only manages
         * tensors whose data is still NULL (cpu_pool.c:290-291), so they are never recycled or run in place. */
        for (int k = 0; k < num_want; k++)
            if (out_bufs && want_ids[k] == L->output)
            {
                set_tensor_shape(ot, dout->dims, 4);
                if (set_tensor_buffer(ot, out_bufs[k], dout->dims[0] * dout->dims[1] * dout->dims[2] * dout->dims[3]) < 0) goto fail;
            }
        tt[L->output] = ot;
        out_node_name[L->output] = strdup(name);

        void* pm = node->op.param_mem;
        switch (L->op)
        {
        case TB200_OP_CONV:
        {
            struct conv_param* p = (struct conv_param*)pm;
            p->kernel_h = L->kernel_h, p->kernel_w = L->kernel_w, p->stride_h = L->stride_h, p->stride_w = L->stride_w;
            p->pad_h0 = L->pad_h0, p->pad_h1 = L->pad_h1, p->pad_w0 = L->pad_w0, p->pad_w1 = L->pad_w1;
            p->dilation_h = L->dilation_h, p->dilation_w = L->dilation_w;
            p->input_channel = din->dims[1], p->output_channel = dout->dims[1], p->group = L->group;
            p->activation = L->activation;
            break;
        }
        case TB200_OP_FC: ((struct fc_param*)pm)->num_output = dout->dims[1]; break;
        case TB200_OP_POOL:
        {
            struct pool_param* p = (struct pool_param*)pm;
            p->pool_method = L->pool_method, p->global = L->pool_global, p->caffe_flavor = L->caffe_flavor;
            p->kernel_h = L->kernel_h, p->kernel_w = L->kernel_w, p->stride_h = L->stride_h, p->stride_w = L->stride_w;
            p->pad_h0 = L->pad_h0, p->pad_h1 = L->pad_h1, p->pad_w0 = L->pad_w0, p->pad_w1 = L->pad_w1;
            /* the *_org values infer_shape starts from (pooling.c:68-92): caffe_flavor 2 splits pad_org into (org/2, org-org/2) */
            p->pad_h0_org = p->pad_h1_org = (L->caffe_flavor == 2) ? L->pad_h0 + L->pad_h1 : L->pad_h0;
            p->pad_w0_org = p->pad_w1_org = (L->caffe_flavor == 2) ? L->pad_w0 + L->pad_w1 : L->pad_w0;
            break;
        }
        case TB200_OP_RELU: ((struct relu_param*)pm)->negative_slope = L->negative_slope; break;
        case TB200_OP_ELTWISE:
        {
            struct eltwise_param* p = (struct eltwise_param*)pm;
            p->type = L->elt_type, p->caffe_flavor = 1;
            break;
        }
        case TB200_OP_CONCAT: ((struct concat_param*)pm)->axis = L->axis; break;
        case TB200_OP_UPSAMPLE: ((struct upsample_param*)pm)->scale = (float)L->up_scale; break;
        case TB200_OP_SOFTMAX: ((struct softmax_param*)pm)->axis = L->axis; break;
        case TB200_OP_RESHAPE: ((struct flatten_param*)pm)->axis = 1, ((struct flatten_param*)pm)->end_axis = 3; break;
        default: break;
        }
    }

    {
        const char** names = (const char**)calloc(num_inputs + num_want, sizeof(char*));
        for (int i = 0; i < num_inputs; i++) names[i] = out_node_name[input_ids[i]];
        if (set_graph_input_node(graph, names, num_inputs) < 0) goto fail;
        for (int i = 0; i < num_want; i++)
        {
            if (!out_node_name[want_ids[i]]) goto fail;
            names[i] = out_node_name[want_ids[i]];
        }
        if (set_graph_output_node(graph, names, num_want) < 0) goto fail;
        free(names);
    }

    B->graph = graph, B->ctx = ctx, B->tt = tt, B->out_node_name = out_node_name, B->precision = precision;
    return 0;
fail:
    B->graph = graph, B->ctx = ctx, B->tt = tt, B->out_node_name = out_node_name;
    free_built(B, num_tensors);
    return rc;
}

/* prerun_graph_multithread() ends with set_cpu_affine(mask) (source/api/c_api.c:540-547), which pins every OpenMP
 * thread of the process to CPUs 0..omp_get_max_threads()-1.  Several reference processes on one host would then all
 * sit on the same cores.  When REF_SHIM_CPUS="3,4,5" is set, the harness re-pins this process' OpenMP team to that
 * CPU list after prerun (process placement only; the reference code is untouched). */
static void repin_from_env(int num_thread)
{
    const char* cs = getenv("REF_SHIM_CPUS");
    if (!cs || !*cs) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (const char* p = cs; *p;)
    {
        char* e;
        long c = strtol(p, &e, 10);
        if (e == p) break;
        if (c >= 0 && c < CPU_SETSIZE) { CPU_SET((int)c, &set); n++; }
        p = (*e == ',') ? e + 1 : e;
        if (*e && *e != ',') break;
    }
    if (!n) return;
    if (num_thread < 1) num_thread = 1;
#pragma omp parallel num_threads(num_thread)
    {
        sched_setaffinity(0, sizeof(set), &set);
    }
}

/* Returns 0 on success.  `want_ids` tensors (layer outputs) are marked graph outputs and returned in out_bufs.
 * ms_stats[0]=min, [1]=avg over `loops` timed run_graph calls (after `warmup`). */
SHIM_API int ref_shim_run(const tb200_tensor_desc* tensors, int num_tensors, const tb200_layer_desc* layers,
                          int num_layers, const int* input_ids, int num_inputs, const int* want_ids, int num_want,
                          const void* const* in_bufs, void* const* out_bufs, const char* device_name,
                          int num_thread, int warmup, int loops, double* ms_stats)
{
    struct built B;
    int rc = build_graph(&B, tensors, num_tensors, layers, num_layers, input_ids, num_inputs, want_ids, num_want, out_bufs, device_name);
    if (rc != 0) return rc;
    graph_t graph = B.graph;
    tensor_t* tt = B.tt;
    int precision = B.precision;
    rc = -1;
    struct options opt;
    opt.num_thread = num_thread;
    opt.cluster = TENGINE_CLUSTER_ALL;
    opt.precision = precision;
    opt.affinity = 0;
    if (prerun_graph_multithread(graph, opt) < 0)
    {
        fprintf(stderr, "ref_shim: prerun failed\n");
        rc = -103;
        goto done;
    }
    repin_from_env(num_thread);
    /* the shapes the reference inferred must be the shapes the caller described */
    for (int i = 0; i < num_tensors; i++)
    {
        if (!tt[i]) continue;
        int dims[8] = {0};
        int nd = get_tensor_shape(tt[i], dims, 8);
        int64_t a = 1, b = 1;
        for (int k = 0; k < nd; k++) a *= dims[k];
        for (int k = 0; k < 4; k++) b *= tensors[i].dims[k];
        if (a != b)
        {
            fprintf(stderr, "ref_shim: tensor %d shape mismatch: reference inferred %d dims [%d %d %d %d]\n", i, nd,
                    dims[0], dims[1], dims[2], dims[3]);
            rc = -104;
            goto done_postrun;
        }
    }
    for (int i = 0; i < num_inputs; i++)
    {
        const tb200_tensor_desc* d = &tensors[input_ids[i]];
        int bytes = d->dims[0] * d->dims[1] * d->dims[2] * d->dims[3];
        if (set_tensor_buffer(tt[input_ids[i]], (void*)in_bufs[i], bytes) < 0) goto done_postrun;
    }
    for (int i = 0; i < warmup; i++)
        if (run_graph(graph, 1) < 0) { rc = -105; goto done_postrun; }
    {
        /* Fleet mode (bench.py's CPU arm): REF_SHIM_WINDOW="<start epoch ms> <end epoch ms>" -- every worker process waits for
         * the common start, then counts the run_graph() calls it COMPLETES before the common end.  Fleet throughput =
         * sum of counts / window, which a starved or late worker lowers by its own share only. */
        double w0 = 0, w1 = 0;
        const char* win = getenv("REF_SHIM_WINDOW");
        const int fleet = win && sscanf(win, "%lf %lf", &w0, &w1) == 2 && w1 > w0;
        double mn = 1e30, sum = 0;
        int done = 0;
        if (fleet)
            while (now_ms() < w0) usleep(200);
        const double t_begin = now_ms();
        for (int i = 0; fleet ? 1 : (i < loops); i++)
        {
            double t0 = now_ms();
            if (fleet && t0 >= w1) break;
            if (run_graph(graph, 1) < 0) { rc = -105; goto done_postrun; }
            double t1 = now_ms();
            if (fleet && t1 > w1) break; /* finished after the window closed: not counted */
            double dt = t1 - t0;
            if (dt < mn) mn = dt;
            sum += dt;
            done++;
        }
        if (ms_stats) ms_stats[0] = mn, ms_stats[1] = done ? sum / done : 0, ms_stats[2] = done, ms_stats[3] = now_ms() - t_begin;
    }
    for (int i = 0; i < num_want; i++)
    {
        const tb200_tensor_desc* d = &tensors[want_ids[i]];
        size_t bytes = (size_t)d->dims[0] * d->dims[1] * d->dims[2] * d->dims[3];
        void* p = get_tensor_buffer(tt[want_ids[i]]);
        if (!p) { rc = -106; goto done_postrun; }
        if (p != out_bufs[i]) memcpy(out_bufs[i], p, bytes);
    }
    rc = 0;
done_postrun:
    postrun_graph(graph);
done:
    free_built(&B, num_tensors);
    return rc;
}

/* Write the graph as a tmfile with the reference's own writer (tools/save_graph/save_graph.cpp:400): the int8/uint8
 * models the UNMODIFIED tm_classification_int8/uint8 and tm_benchmark binaries are then run on. */
extern int ref_shim_save_graph_cxx(void* graph, const char* fname);
extern int infer_ir_graph_shape(struct graph* graph);
SHIM_API int ref_shim_save_tmfile(const tb200_tensor_desc* tensors, int num_tensors, const tb200_layer_desc* layers,
                                  int num_layers, const int* input_ids, int num_inputs, const int* output_ids,
                                  int num_outputs, const char* fname)
{
    struct built B;
    int rc = build_graph(&B, tensors, num_tensors, layers, num_layers, input_ids, num_inputs, output_ids, num_outputs, NULL, NULL);
    if (rc != 0) return rc;
    if (infer_ir_graph_shape((struct graph*)B.graph) != 0) rc = -110;
    else
    {
        /* The tmfile's pad fields are what the loader turns into pad_*_org (serializer/tmfile/op/tm2_pool.c:56-64), from which
         * infer_shape derives the real pads again (operator/prototype/pooling.c:68-92): store the *_org values, as a
         * converter-written model does. */
        struct graph* g = (struct graph*)B.graph;
        for (int i = 0; i < g->node_num; i++)
            if (g->node_list[i]->op.type == OP_POOL)
            {
                struct pool_param* p = (struct pool_param*)g->node_list[i]->op.param_mem;
                p->pad_h0 = p->pad_h0_org, p->pad_h1 = p->pad_h1_org, p->pad_w0 = p->pad_w0_org, p->pad_w1 = p->pad_w1_org;
            }
        rc = ref_shim_save_graph_cxx(B.graph, fname);
    }
    free_built(&B, num_tensors);
    return rc;
}

SHIM_API const char* ref_shim_version(void) { return get_tengine_version(); }

/* One fp32 convolution on the reference CPU device (TENGINE_MODE_FP32): the fp32 members of the path -- Winograd F(4,3)
 * (wino_conv_hcl, selected by winograd_support(), conv_kernel_x86.c:1896-1915), im2col + sgemm_fp, fp32 depthwise
 * (conv_dw_kernel_x86.c).  Built the way tests/op/test_op.h:619-664 builds 1-op graphs.  x [n,c,h,w], wt [oc,c/group,kh,kw]. */
SHIM_API int ref_shim_conv_f32(int n, int c, int h, int w, int oc, int kh, int kw, int stride, int pad, int group, int activation, const float* x,
                               const float* wt, const float* bias, float* y, int num_thread)
{
    if (!g_inited)
    {
        if (init_tengine() != 0) return -100;
        g_inited = 1;
    }
    int rc = -1;
    graph_t graph = create_graph(NULL, NULL, NULL);
    if (!graph) return -102;
    node_t in_node = create_graph_node(graph, "in", "InputOp");
    tensor_t in_t = create_graph_tensor(graph, "in", TENGINE_DT_FP32);
    set_node_output_tensor(in_node, 0, in_t, TENSOR_TYPE_INPUT);
    int dims[4] = {n, c, h, w};
    set_tensor_shape(in_t, dims, 4);
    int wdims[4] = {oc, c / group, kh, kw};
    node_t wn = create_graph_node(graph, "w", "Const");
    tensor_t wtn = create_graph_tensor(graph, "w", TENGINE_DT_FP32);
    set_node_output_tensor(wn, 0, wtn, TENSOR_TYPE_CONST);
    set_tensor_shape(wtn, wdims, 4);
    set_tensor_buffer(wtn, (void*)wt, oc * (c / group) * kh * kw * 4);
    tensor_t btn = NULL;
    if (bias)
    {
        node_t bn = create_graph_node(graph, "b", "Const");
        btn = create_graph_tensor(graph, "b", TENGINE_DT_FP32);
        set_node_output_tensor(bn, 0, btn, TENSOR_TYPE_CONST);
        int bd[1] = {oc};
        set_tensor_shape(btn, bd, 1);
        set_tensor_buffer(btn, (void*)bias, oc * 4);
    }
    struct node* node = (struct node*)create_graph_node(graph, "conv", "Convolution");
    tensor_t out_t = create_graph_tensor(graph, "conv", TENGINE_DT_FP32);
    set_node_input_tensor(node, 0, in_t);
    set_node_input_tensor(node, 1, wtn);
    if (btn) set_node_input_tensor(node, 2, btn);
    set_node_output_tensor(node, 0, out_t, TENSOR_TYPE_VAR);
    struct conv_param* p = (struct conv_param*)node->op.param_mem;
    p->kernel_h = kh, p->kernel_w = kw, p->stride_h = p->stride_w = stride, p->pad_h0 = p->pad_h1 = p->pad_w0 = p->pad_w1 = pad;
    p->dilation_h = p->dilation_w = 1, p->input_channel = c, p->output_channel = oc, p->group = group, p->activation = activation;
    const char* in_names[1] = {"in"};
    const char* out_names[1] = {"conv"};
    if (set_graph_input_node(graph, in_names, 1) < 0 || set_graph_output_node(graph, out_names, 1) < 0) goto done;
    set_tensor_buffer(in_t, (void*)x, n * c * h * w * 4);
    struct options opt;
    opt.num_thread = num_thread, opt.cluster = TENGINE_CLUSTER_ALL, opt.precision = TENGINE_MODE_FP32, opt.affinity = 0;
    if (prerun_graph_multithread(graph, opt) < 0) { rc = -103; goto done; }
    if (run_graph(graph, 1) < 0) { rc = -105; postrun_graph(graph); goto done; }
    {
        int od[4];
        get_tensor_shape(out_t, od, 4);
        memcpy(y, get_tensor_buffer(out_t), (size_t)od[0] * od[1] * od[2] * od[3] * 4);
    }
    postrun_graph(graph);
    rc = 0;
done:
    destroy_graph(graph);
    return rc;
}

/* Load a tmfile with the reference's serializer and run it on a named device ("CPU" / NULL, or e.g. "B200") exactly the way
 * tm_benchmark does (create_context + set_context_device + create_graph + set_tensor_shape/buffer + prerun_graph_multithread +
 * run_graph, benchmark/tm_benchmark.cc:60-135), with a caller-supplied NCHW input of `batch` images; copies graph output i
 * into out_bufs[i] (out_bytes[i] bytes available; the real size is written back).  precision: TENGINE_MODE_*.  opt_blob: the
 * device's option struct for set_context_device (may be NULL). */
SHIM_API int ref_shim_run_tmfile(const char* fname, const char* device_name, const void* opt_blob, int opt_size, int precision, const int* in_dims,
                                 const void* in_buf, int num_out, void* const* out_bufs, int64_t* out_bytes, int num_thread, int warmup, int loops,
                                 double* ms_stats)
{
    if (!g_inited)
    {
        if (init_tengine() != 0) return -100;
        g_inited = 1;
    }
    int rc = -1;
    context_t ctx = NULL;
    if (device_name && strcmp(device_name, "CPU") != 0)
    {
        ctx = create_context("shim_ctx", 1);
        if (set_context_device(ctx, device_name, opt_blob, opt_size) < 0)
        {
            fprintf(stderr, "ref_shim: device %s not available\n", device_name);
            destroy_context(ctx);
            return -101;
        }
    }
    graph_t graph = create_graph(ctx, "tengine", fname);
    if (!graph)
    {
        if (ctx) destroy_context(ctx);
        return -102;
    }
    tensor_t it = get_graph_input_tensor(graph, 0, 0);
    int64_t in_bytes = (int64_t)in_dims[0] * in_dims[1] * in_dims[2] * in_dims[3];
    if (!it || set_tensor_shape(it, in_dims, 4) < 0 || set_tensor_buffer(it, (void*)in_buf, (int)in_bytes) < 0) goto done;
    struct options opt;
    opt.num_thread = num_thread, opt.cluster = TENGINE_CLUSTER_ALL, opt.precision = precision, opt.affinity = 0;
    if (prerun_graph_multithread(graph, opt) < 0)
    {
        rc = -103;
        goto done;
    }
    repin_from_env(num_thread);
    for (int i = 0; i < warmup; i++)
        if (run_graph(graph, 1) < 0) { rc = -105; goto done_postrun; }
    {
        double mn = 1e30, sum = 0;
        for (int i = 0; i < loops; i++)
        {
            double t0 = now_ms();
            if (run_graph(graph, 1) < 0) { rc = -105; goto done_postrun; }
            double dt = now_ms() - t0;
            if (dt < mn) mn = dt;
            sum += dt;
        }
        if (ms_stats) ms_stats[0] = mn, ms_stats[1] = loops ? sum / loops : 0;
    }
    if (get_graph_output_node_number(graph) < num_out) { rc = -107; goto done_postrun; }
    for (int i = 0; i < num_out; i++)
    {
        tensor_t ot = get_graph_output_tensor(graph, i, 0);
        const int64_t bytes = ot ? get_tensor_buffer_size(ot) : 0;
        void* p = ot ? get_tensor_buffer(ot) : NULL;
        if (!p || bytes > out_bytes[i]) { rc = -106; goto done_postrun; }
        memcpy(out_bufs[i], p, (size_t)bytes);
        out_bytes[i] = bytes;
    }
    rc = 0;
done_postrun:
    postrun_graph(graph);
done:
    destroy_graph(graph);
    if (ctx) destroy_context(ctx);
    return rc;
}

/* Print op + parameters of every non-const node of a tmfile (used to mirror the reference's benchmark graphs in
 * tengine_b200/workloads.py). */
SHIM_API int ref_shim_describe_tmfile(const char* fname)
{
    if (!g_inited)
    {
        if (init_tengine() != 0) return -100;
        g_inited = 1;
    }
    graph_t graph = create_graph(NULL, "tengine", fname);
    if (!graph) return -1;
    struct graph* g = (struct graph*)graph;
    infer_ir_graph_shape(g);
    for (int i = 0; i < g->node_num; i++)
    {
        struct node* n = g->node_list[i];
        if (n->op.type == OP_CONST) continue;
        struct tensor* o = g->tensor_list[n->output_tensors[0]];
        printf("%d %s out[%d,%d,%d,%d] in:", i, get_node_op(n), o->dims[0], o->dims[1], o->dims[2], o->dims[3]);
        for (int k = 0; k < n->input_num; k++)
        {
            struct tensor* t = g->tensor_list[n->input_tensors[k]];
            if (t->tensor_type != TENSOR_TYPE_CONST) printf(" n%d", t->producer);
        }
        if (n->op.type == OP_CONV)
        {
            struct conv_param* p = (struct conv_param*)n->op.param_mem;
            printf(" k%dx%d s%d,%d p%d,%d,%d,%d d%d g%d act%d oc%d", p->kernel_h, p->kernel_w, p->stride_h, p->stride_w, p->pad_h0, p->pad_h1,
                   p->pad_w0, p->pad_w1, p->dilation_h, p->group, p->activation, p->output_channel);
        }
        else if (n->op.type == OP_POOL)
        {
            struct pool_param* p = (struct pool_param*)n->op.param_mem;
            printf(" method%d k%dx%d s%d,%d p%d,%d,%d,%d porg%d,%d,%d,%d global%d caffe%d", p->pool_method, p->kernel_h, p->kernel_w, p->stride_h,
                   p->stride_w, p->pad_h0, p->pad_h1, p->pad_w0, p->pad_w1, p->pad_h0_org, p->pad_h1_org, p->pad_w0_org, p->pad_w1_org,
                   p->global, p->caffe_flavor);
        }
        else if (n->op.type == OP_RELU)
            printf(" slope%g", ((struct relu_param*)n->op.param_mem)->negative_slope);
        else if (n->op.type == OP_ELTWISE)
            printf(" type%d", ((struct eltwise_param*)n->op.param_mem)->type);
        else if (n->op.type == OP_UPSAMPLE)
            printf(" scale%g", ((struct upsample_param*)n->op.param_mem)->scale);
        else if (n->op.type == OP_CONCAT)
            printf(" axis%d", ((struct concat_param*)n->op.param_mem)->axis);
        printf("\n");
    }
    destroy_graph(graph);
    return 0;
}

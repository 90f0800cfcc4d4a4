// yolo_detect.cu -- detection post-processing on the device: YOLO region decode, score threshold, sort by
// score and greedy NMS, straight from the graph's quantised output tensors in HBM.  Only the kept boxes travel to the host
// (YOLOv3-tiny at batch 128: 27.6 MB of raw head tensors stay on the GPU).
//
// Restates the application code of examples/tm_yolov3_tiny_uint8.cpp: dequantisation (:464-478), generate_proposals (:176-250),
// qsort_descent_inplace (:57-100), nms_sorted_bboxes (:102-132).  Every per-element function of a byte there -- sigmoid, exp --
// is a 256-entry table built on the host with the same float arithmetic (engine.cu build_yolo_tables), so the device only
// looks up and multiplies; sorting and NMS replay the example's algorithms literally (including its quicksort's tie order),
// one CTA per image.
#include "common.cuh"
#include "kernels.h"

namespace tb200 {

// ---- decode: one thread per (image, cell, anchor) of one head ------------------------------------------------------------------
__global__ void __launch_bounds__(256) yolo_decode_kernel(const uint8_t* __restrict__ t, int cp, int H, int W, int n_img, int anchors_n, int classes,
                                                          const float* __restrict__ sig,  // [256]: sigmoid(dequantised byte), as a float
                                                          const double* __restrict__ ex,  // [256]: the example's float exp(dequantised byte), held as a double
                                                          float stride, float a0w, float a0h, float a1w, float a1h, float a2w, float a2h, float thr,
                                                          YoloCand* __restrict__ cand, int* __restrict__ count, int max_cand, unsigned key_base, bool is_u8)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_img * H * W * anchors_n;
    if (idx >= total) return;
    const int anchor = (int)(idx % anchors_n);
    const long long cell = idx / anchors_n;
    const int w = (int)(cell % W), h = (int)((cell / W) % H), n = (int)(cell / ((long long)W * H));
    const uint8_t* p = t + (size_t)cell * cp + (size_t)anchor * (classes + 5);
    // class arg-max: dequantisation is increasing in the byte, so the first maximal byte is the reference's `score > class_score`
    int best = 0, bestv = -1000;
    for (int s = 0; s < classes; s++)
    {
        const int v = is_u8 ? (int)p[5 + s] : (int)(int8_t)p[5 + s];
        if (v > bestv) bestv = v, best = s;
    }
    const float final_score = __fmul_rn(sig[p[4]], sig[p[5 + best]]);
    if (!(final_score >= thr)) return;
    const float dx = sig[p[0]], dy = sig[p[1]];
    const float aw = anchor == 0 ? a0w : (anchor == 1 ? a1w : a2w), ah = anchor == 0 ? a0h : (anchor == 1 ? a1h : a2h);
    const float pred_x = __fmul_rn(__fadd_rn((float)w, dx), stride), pred_y = __fmul_rn(__fadd_rn((float)h, dy), stride);
    // `float pred_w = exp(dw) * anchor_w;` is a float product (exp resolves to the float overload there): the double product of two
    // float values is exact, so narrowing it rounds once, exactly like the float multiply
    const float pred_w = (float)__dmul_rn(ex[p[2]], (double)aw), pred_h = (float)__dmul_rn(ex[p[3]], (double)ah);
    const float x0 = __fsub_rn(pred_x, __fmul_rn(pred_w, 0.5f)), y0 = __fsub_rn(pred_y, __fmul_rn(pred_h, 0.5f));
    const float x1 = __fadd_rn(pred_x, __fmul_rn(pred_w, 0.5f)), y1 = __fadd_rn(pred_y, __fmul_rn(pred_h, 0.5f));
    const int slot = atomicAdd(count + n, 1);
    if (slot >= max_cand) return; // counted, reported as overflow by the host
    YoloCand c;
    c.x = x0, c.y = y0, c.w = __fsub_rn(x1, x0), c.h = __fsub_rn(y1, y0), c.prob = final_score, c.label = best;
    c.key = key_base + (unsigned)((h * W + w) * anchors_n + anchor); // position in the reference's proposal list
    cand[(size_t)n * max_cand + slot] = c;
}

// ---- per image: restore the proposal order, the example's quicksort, greedy NMS ---------------------------------------------------
__device__ void yolo_qsort(YoloCand* v, int left, int right)
{
    // qsort_descent_inplace (tm_yolov3_tiny_uint8.cpp:57-92) with an explicit stack (the recursion's two halves are independent)
    int stack[64][2];
    int sp = 0;
    stack[sp][0] = left, stack[sp][1] = right, sp++;
    while (sp > 0)
    {
        sp--;
        const int l = stack[sp][0], r = stack[sp][1];
        int i = l, j = r;
        const float p = v[(l + r) / 2].prob;
        while (i <= j)
        {
            while (v[i].prob > p) i++;
            while (v[j].prob < p) j--;
            if (i <= j)
            {
                const YoloCand tmp = v[i];
                v[i] = v[j], v[j] = tmp;
                i++, j--;
            }
        }
        // push the larger part first so that the stack depth stays logarithmic; the parts are disjoint, so the order they are
        // sorted in does not change the result
        const bool left_ok = l < j, right_ok = i < r;
        if (left_ok && right_ok)
        {
            if (j - l > r - i)
                stack[sp][0] = l, stack[sp][1] = j, sp++, stack[sp][0] = i, stack[sp][1] = r, sp++;
            else
                stack[sp][0] = i, stack[sp][1] = r, sp++, stack[sp][0] = l, stack[sp][1] = j, sp++;
        }
        else if (left_ok)
            stack[sp][0] = l, stack[sp][1] = j, sp++;
        else if (right_ok)
            stack[sp][0] = i, stack[sp][1] = r, sp++;
    }
}

__global__ void __launch_bounds__(256) yolo_nms_kernel(const YoloCand* __restrict__ cand, const int* __restrict__ count, int max_cand, float nms_thr,
                                                       YoloCand* __restrict__ sorted, // scratch [n_img][max_cand]
                                                       YoloDet* __restrict__ out, int max_out, int* __restrict__ out_count)
{
    const int n = blockIdx.x;
    int m = count[n];
    __shared__ int s_picked[1024];
    __shared__ int s_np, s_keep;
    if (m > max_cand)
    {
        if (threadIdx.x == 0) out_count[n] = -m; // overflow: the caller must raise max_candidates
        return;
    }
    const YoloCand* src = cand + (size_t)n * max_cand;
    YoloCand* v = sorted + (size_t)n * max_cand;
    // (1) proposal order of the reference = ascending key (keys are unique): rank sort
    for (int i = threadIdx.x; i < m; i += blockDim.x)
    {
        const unsigned k = src[i].key;
        int rank = 0;
        for (int j = 0; j < m; j++) rank += src[j].key < k;
        v[rank] = src[i];
    }
    __syncthreads();
    // (2) the example's quicksort by score, literally (one thread: its tie order is part of the result)
    if (threadIdx.x == 0)
    {
        if (m > 0) yolo_qsort(v, 0, m - 1);
        s_np = 0;
    }
    __syncthreads();
    // (3) nms_sorted_bboxes (:102-132): candidate i is kept unless its IoU with an already kept box exceeds the threshold
    for (int i = 0; i < m; i++)
    {
        if (threadIdx.x == 0) s_keep = 1;
        __syncthreads();
        const YoloCand a = v[i];
        const float area_a = __fmul_rn(a.w, a.h);
        const int np = s_np;
        for (int j = threadIdx.x; j < np; j += blockDim.x)
        {
            const YoloCand b = v[s_picked[j]];
            // cv::Rect_<float> operator& and area()
            const float x1 = fmaxf(a.x, b.x), y1 = fmaxf(a.y, b.y);
            const float iw = __fsub_rn(fminf(__fadd_rn(a.x, a.w), __fadd_rn(b.x, b.w)), x1), ih = __fsub_rn(fminf(__fadd_rn(a.y, a.h), __fadd_rn(b.y, b.h)), y1);
            const float inter = (iw <= 0.f || ih <= 0.f) ? 0.f : __fmul_rn(iw, ih);
            const float uni = __fsub_rn(__fadd_rn(area_a, __fmul_rn(b.w, b.h)), inter);
            if (__fdiv_rn(inter, uni) > nms_thr) s_keep = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_keep)
        {
            if (s_np < 1024) s_picked[s_np] = i;
            s_np++;
        }
        __syncthreads();
        if (s_np > 1024) break; // more kept boxes than the list holds: reported below
    }
    __syncthreads();
    const int np = s_np;
    if (threadIdx.x == 0) out_count[n] = (np > 1024 || np > max_out) ? -np : np;
    for (int j = threadIdx.x; j < np && j < max_out && j < 1024; j += blockDim.x)
    {
        const YoloCand c = v[s_picked[j]];
        YoloDet d;
        d.x = c.x, d.y = c.y, d.w = c.w, d.h = c.h, d.prob = c.prob, d.label = c.label;
        out[(size_t)n * max_out + j] = d;
    }
}

cudaError_t launch_yolo_decode(const void* tensor, int cp, int h, int w, int n_img, int anchors_n, int classes, const float* sig, const double* ex, float stride,
                               const float* anchors6, float thr, YoloCand* cand, int* count, int max_cand, unsigned key_base, bool is_u8, cudaStream_t st)
{
    const long long total = (long long)n_img * h * w * anchors_n;
    if (total <= 0 || total >= (1ll << 31) || anchors_n > 3) return cudaErrorInvalidValue;
    yolo_decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint8_t*)tensor, cp, h, w, n_img, anchors_n, classes, sig, ex, stride, anchors6[0],
                                                                      anchors6[1], anchors6[2], anchors6[3], anchors6[4], anchors6[5], thr, cand, count, max_cand,
                                                                      key_base, is_u8);
    return cudaGetLastError();
}

cudaError_t launch_yolo_nms(const YoloCand* cand, const int* count, int n_img, int max_cand, float nms_thr, YoloCand* sorted, YoloDet* out, int max_out,
                            int* out_count, cudaStream_t st)
{
    yolo_nms_kernel<<<n_img, 256, 0, st>>>(cand, count, max_cand, nms_thr, sorted, out, max_out, out_count);
    return cudaGetLastError();
}

} // namespace tb200

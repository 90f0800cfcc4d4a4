// TEST INFRASTRUCTURE (oracle): the detection post-processing of the UNMODIFIED reference example, compiled from the source where it
// lies (examples/tm_yolov3_tiny_uint8.cpp; its main() is renamed, the C++ OpenCV it needs is replaced by oracle/cvstub) and exported
// through one C function, so that oracle/yolo_post.py -- the checker of tb200_graph_yolo_detect -- is pinned against the example's
// own code instead of a reading of it.  Built by oracle/build_ref.py into oracle/_ref/libyolo_example.so.
#define main tm_yolov3_tiny_uint8_example_main
#include "tm_yolov3_tiny_uint8.cpp"
#undef main

// p32 / p16: the two dequantised head tensors of ONE image, [255][13][13] and [255][26][26] floats, exactly what the example's
// main() builds at :464-478.  The call sequence below is main():480-497.  Returns the number of kept boxes, written as
// (x, y, w, h, prob, label) rows.
extern "C" int yolo_example_postprocess(const float* p32, const float* p16, float prob_threshold, float nms_threshold, float* out6, int max_out)
{
    std::vector<Object> proposals, objects16, objects32;
    generate_proposals(32, p32, prob_threshold, objects32);
    proposals.insert(proposals.end(), objects32.begin(), objects32.end());
    generate_proposals(16, p16, prob_threshold, objects16);
    proposals.insert(proposals.end(), objects16.begin(), objects16.end());
    qsort_descent_inplace(proposals);
    std::vector<int> picked;
    nms_sorted_bboxes(proposals, picked, nms_threshold);
    int n = 0;
    for (size_t i = 0; i < picked.size() && n < max_out; i++, n++)
    {
        const Object& o = proposals[picked[i]];
        float* r = out6 + 6 * n;
        r[0] = o.rect.x, r[1] = o.rect.y, r[2] = o.rect.width, r[3] = o.rect.height, r[4] = o.prob, r[5] = (float)o.label;
    }
    return (int)picked.size();
}

extern "C" float yolo_example_sigmoid(float x) { return sigmoid(x); }

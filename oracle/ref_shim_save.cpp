// ref_shim_save.cpp -- C bridge to the reference's tmfile writer (tools/save_graph/save_graph.cpp).  TEST INFRASTRUCTURE.
#include "tengine/c_api.h"
#include "save_graph.hpp"
extern "C" int ref_shim_save_graph_cxx(void* graph, const char* fname) { return save_graph((graph_t)graph, fname) ? 0 : -111; }

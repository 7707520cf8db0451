/* bvh8_build.h -- host conversion of the reference's acceleration structures to the device layout (bvh8.h). */
#ifndef IMBA_BVH8_BUILD_H
#define IMBA_BVH8_BUILD_H
#include <map>
#include <string>
#include <vector>
#include "bvh8.h"
#include "../../include/imba_b200.h"

struct Bvh8 {
    std::vector<Node8>  nodes;
    std::vector<Tri48>  tris;
    std::vector<Inst64> insts;
    uint32_t top_root = 0;   // Node8 index of the top-level root
    int      max_depth = 0;  // deepest root-to-leaf path (top level + mesh level)
    int      ref_node_count = 0; // nodes of the reference structure that were visited
    // single instance with the identity transform (every scene of the reference's test set except cornell_instances):
    // traversal may start in the mesh tree directly, the result is bit-identical (1*x + 0 is exact)
    bool     flat = false; uint32_t flat_root = 0; int32_t flat_inst = 0;
};

/* layout: IMBA_ACCEL_MBVH4 (nodes = imba_node4[]) or IMBA_ACCEL_BVH2 (nodes = imba_node2[]); all pointers are host
 * pointers.  Returns false and sets err on malformed input (the reference has no such check: it asserts). */
bool imba_build_bvh8(int layout, int root, const void* nodes, const imba_vec4* tris,
                     const imba_instance_node* inst_nodes, Bvh8& out, std::string& err);

/* Top level only (the mesh trees were built on the device): see bvh8_build.cpp. */
bool imba_build_bvh8_top(int layout, int root, const void* nodes, const imba_instance_node* inst_nodes,
                         const std::map<int, uint32_t>& mesh_roots, uint32_t node_offset, Bvh8& out, std::string& err);

#endif

/* ORACLE / TEST INFRASTRUCTURE. Stand-in for AnyDSL's generated traversal_gpu.h. */
#include "traversal_types.inc"
struct BBox { float lo_x, lo_y, lo_z, hi_x, hi_y, hi_z; };
struct Node { BBox left_bb; BBox right_bb; int left; int right; };
extern "C" {
void intersect_gpu_masked_instanced(int root, Node* nodes, InstanceNode* instances, Vec4* tris, Ray* rays, Hit* hits,
                                    int* indices, Vec2* texcoords, TransparencyMask* masks, char* mask_buffer, int count);
void occluded_gpu_masked_instanced(int root, Node* nodes, InstanceNode* instances, Vec4* tris, Ray* rays, Hit* hits,
                                   int* indices, Vec2* texcoords, TransparencyMask* masks, char* mask_buffer, int count);
}

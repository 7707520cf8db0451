/* ORACLE / TEST INFRASTRUCTURE. Stand-in for AnyDSL's generated traversal_cpu.h. */
#include "traversal_types.inc"
struct Node { float min_x[4], min_y[4], min_z[4], max_x[4], max_y[4], max_z[4]; int children[4]; };
extern "C" {
void intersect_cpu_masked_instanced(int root, Node* nodes, InstanceNode* instances, Vec4* tris, Ray* rays, Hit* hits,
                                    int* indices, Vec2* texcoords, TransparencyMask* masks, char* mask_buffer, int count);
void occluded_cpu_masked_instanced(int root, Node* nodes, InstanceNode* instances, Vec4* tris, Ray* rays, Hit* hits,
                                   int* indices, Vec2* texcoords, TransparencyMask* masks, char* mask_buffer, int count);
}

/* bvh_cache.cpp -- reader / writer of the reference's per-mesh `.bvh` acceleration-structure cache
 * (loaders/load_bvh.cpp:12-43, 144-302; written automatically next to the mesh when a scene file has an `accel` line,
 * render/scene.cpp:67).
 *
 *   file   = u32 magic 0x313F1A57, then blocks
 *   block  = u64 size (bytes after this field), u32 type, Header{u32 node_count, u32 prim_count}, nodes[], prims[]
 *   type 1 = BVH  (traversal_gpu::Node, 56 B; prims = Vec4, 3 per triangle)      type 2 = MBVH (traversal_cpu::Node, 112 B;
 *            prims = Vec4, 13 per packet + sentinels); indices inside a block are file-local (node 0 = mesh root, triangle ids
 *            start at 0), load_bvh.cpp:199-224
 *   type 3 = (new) the compressed 8-wide layout of this library: Node8 (80 B) + Tri48 (48 B, incl. nx / ny), csrc/bvh8.h, so that a cache
 *            written once skips the conversion; unknown block types are skipped by the reference's reader (locate_block).
 */
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "api_common.h"

namespace {
constexpr uint32_t MAGIC = 0x313F1A57;
struct Block { uint32_t type, node_count, prim_count; std::vector<unsigned char> nodes, prims; };
size_t node_bytes(uint32_t type) { return type == 1 ? 56 : type == 2 ? 112 : type == 3 ? 80 : 0; }
size_t prim_bytes(uint32_t type) { return type == 3 ? 48 : 16; }
}

struct imba_bvh_file { std::vector<Block> blocks; };

extern "C" void imba_bvh_cache_close(imba_bvh_file* f) { delete f; }

extern "C" imba_bvh_file* imba_bvh_cache_open(const char* path) {
    FILE* in = path ? fopen(path, "rb") : nullptr;
    if (!in) { imba_set_error("cannot open %s", path ? path : "(null)"); return nullptr; }
    uint32_t magic = 0;
    if (fread(&magic, 4, 1, in) != 1 || magic != MAGIC) { fclose(in); imba_set_error("%s is not a .bvh cache (bad magic)", path); return nullptr; }
    fseek(in, 0, SEEK_END); const uint64_t file_size = (uint64_t)ftell(in); fseek(in, 4, SEEK_SET);
    imba_bvh_file* f = new imba_bvh_file;
    try {
    for (;;) {
        uint64_t size = 0; uint32_t type = 0;
        if (fread(&size, 8, 1, in) != 1) break;
        if (fread(&type, 4, 1, in) != 1 || size < 4) { imba_set_error("truncated block header in %s", path); delete f; fclose(in); return nullptr; }
        const uint64_t payload = size - 4;
        const size_t nb = node_bytes(type);
        if (nb == 0) { if (fseek(in, (long)payload, SEEK_CUR) != 0) break; continue; } // unknown block: skip
        Block b; b.type = type;
        uint32_t h[2];
        if (payload < 8 || fread(h, 4, 2, in) != 2) { imba_set_error("truncated block in %s", path); delete f; fclose(in); return nullptr; }
        b.node_count = h[0]; b.prim_count = h[1];
        const uint64_t want = 8 + (uint64_t)nb * h[0] + (uint64_t)prim_bytes(type) * h[1];
        if (want != payload) { imba_set_error("block of type %u in %s has inconsistent size", type, path); delete f; fclose(in); return nullptr; }
        if (payload > file_size) { imba_set_error("block of type %u in %s is larger than the file", type, path); delete f; fclose(in); return nullptr; }   // untrusted counts never size an allocation beyond the file
        b.nodes.resize(nb * h[0]); b.prims.resize(prim_bytes(type) * h[1]);
        if ((b.nodes.size() && fread(b.nodes.data(), 1, b.nodes.size(), in) != b.nodes.size()) ||
            (b.prims.size() && fread(b.prims.data(), 1, b.prims.size(), in) != b.prims.size())) {
            imba_set_error("truncated block data in %s", path); delete f; fclose(in); return nullptr;
        }
        f->blocks.push_back(std::move(b));
    }
    } catch (const std::bad_alloc&) {   // no exception crosses the C ABI
        imba_set_error("out of memory reading %s", path); delete f; fclose(in); return nullptr;
    }
    fclose(in);
    return f;
}

/* What the reference's store_accel_* looks at before appending (load_bvh.cpp:144-157): the magic, and the block headers for a
 * block of the same type -- without loading or validating any payload.  Returns -1 no such file, -2 not a .bvh cache, 1 a block of
 * `type` is present, 0 otherwise. */
static int scan_headers(const char* path, uint32_t type) {
    FILE* in = fopen(path, "rb");
    if (!in) return -1;
    uint32_t magic = 0;
    if (fread(&magic, 4, 1, in) != 1 || magic != MAGIC) { fclose(in); return -2; }
    int found = 0;
    for (;;) {
        uint64_t size = 0; uint32_t t = 0;
        if (fread(&size, 8, 1, in) != 1 || fread(&t, 4, 1, in) != 1 || size < 4) break;   // end of file or a damaged tail: stop scanning
        if (t == type) { found = 1; break; }
        if (fseek(in, (long)(size - 4), SEEK_CUR) != 0) break;
    }
    fclose(in);
    return found;
}

extern "C" int imba_bvh_cache_block(const imba_bvh_file* f, int type, const void** nodes, uint32_t* node_count, const void** prims, uint32_t* prim_count) {
    if (!f) { imba_set_error("null argument"); return IMBA_ERR_INVALID; }
    for (const Block& b : f->blocks) if ((int)b.type == type) {
        if (nodes) *nodes = b.nodes.data();
        if (node_count) *node_count = b.node_count;
        if (prims) *prims = b.prims.data();
        if (prim_count) *prim_count = b.prim_count;
        return IMBA_OK;
    }
    imba_set_error("no block of type %d", type);
    return IMBA_ERR_INVALID;
}

/* Appends one block (creating the file with its magic if needed).  Like store_accel_* (load_bvh.cpp:144-157) it refuses
 * to add a second block of the same type. */
extern "C" int imba_bvh_cache_append(const char* path, int type, const void* nodes, uint32_t node_count, const void* prims, uint32_t prim_count) {
    const size_t nb = node_bytes((uint32_t)type);
    if (!path || nb == 0 || (node_count && !nodes) || (prim_count && !prims)) { imba_set_error("invalid argument"); return IMBA_ERR_INVALID; }
    // existence is decided from the magic alone, as the reference does: a file whose later blocks are damaged is appended to, never
    // truncated (its good blocks -- e.g. the BVH the reference's own loader cached -- stay); a file that is not a cache is refused
    const int state = scan_headers(path, (uint32_t)type);
    if (state == -2) { imba_set_error("%s exists and is not a .bvh cache: not overwritten", path); return IMBA_ERR_INVALID; }
    if (state == 1) { imba_set_error("%s already contains a block of type %d", path, type); return IMBA_ERR_INVALID; }
    const bool exists = state >= 0;
    FILE* out = fopen(path, exists ? "ab" : "wb");
    if (!out) { imba_set_error("cannot write %s", path); return IMBA_ERR_INVALID; }
    if (!exists) fwrite(&MAGIC, 4, 1, out);
    const uint32_t t = (uint32_t)type, h[2] = { node_count, prim_count };
    const uint64_t size = 4 + 8 + (uint64_t)nb * node_count + (uint64_t)prim_bytes(t) * prim_count;
    fwrite(&size, 8, 1, out); fwrite(&t, 4, 1, out); fwrite(h, 4, 2, out);
    if (node_count) fwrite(nodes, nb, node_count, out);
    if (prim_count) fwrite(prims, prim_bytes(t), prim_count, out);
    const bool ok = !ferror(out);
    fclose(out);
    if (!ok) { imba_set_error("write error on %s", path); return IMBA_ERR_INVALID; }
    return IMBA_OK;
}

// bvh_sah_gpu.cu — the FAST traversal tree (binned SAH, <= 4 primitives per leaf) built on the GPU.
//
// Level-synchronous top-down build of the tree specified in sah_split.h; the result equals the host builder's
// (bvh_sah.cpp) node for node and bit for bit, which tests/test_gpu_bvh_build.py checks on every fixture scene and on the
// 868 334-triangle stand-in.  One level = one pass over the primitive permutation:
//
//   bin_kernel        every primitive adds its box to the bin of its node on each axis with a positive centroid extent
//                     (atomicMin / atomicMax on order-preserving integer keys + an atomic count: exact, order independent)
//   decide_kernel     one thread per node evaluates sah_decide on those bins (the code the host runs)
//   scan              exclusive prefix sums: "goes left" flags over the permutation (stable partition), "splits" flags over
//                     the level's nodes (child numbering)
//   emit_kernel       one thread per node: leaf reference or child ranges, written into the parent's record
//   partition_kernel  every primitive moves to its place in the child's range and adds its box and centroid to the child
//
// Nodes are created in breadth-first order; three more sweeps over the levels (subtree sizes bottom-up, pre-order index
// top-down, emit) renumber them depth-first, the layout the traversal kernels were tuned on.  Everything is data parallel
// over primitives or nodes; the host only reads one counter per level to size the next launch.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "bvh_sah.h"
#include "sah_split.h"

namespace cge {
namespace {

constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kKeyEmptyLo = 0xff7fffffu; // sah_key(+FLT_MAX)
constexpr uint32_t kKeyEmptyHi = 0x00800000u; // sah_key(-FLT_MAX)

struct PrimBox { // 40 bytes per primitive, by global primitive id
    float lo[3], hi[3], c[3];
    uint32_t pad;
};

// A node of the level being processed (or of the next one while it is filled in).
struct LevelNode {
    uint32_t beg, end;     // range of the permutation
    uint32_t lo[3], hi[3]; // bounds of the primitives, as keys
    uint32_t clo[3], chi[3]; // centroid bounds, as keys
    uint32_t parent;       // breadth-first index of the parent inner node, kNone for the root
    uint32_t side;         // 0 = left child, 1 = right child
};

struct NodeBins { // 3 axes x 16 bins: box keys + count
    uint32_t lo[3][kSahBins][3], hi[3][kSahBins][3];
    uint32_t count[3][kSahBins];
};

struct Decision {
    int kind, axis, bin;
    float lo, scale;
};

// Inner node in breadth-first order.  Child reference: packed leaf (bit 31) or breadth-first index of an inner node.
struct BfsNode {
    float box[2][6]; // [side]: lo.xyz, hi.xyz of the child
    uint32_t child[2];
};

__device__ __forceinline__ SahBox box_from_keys(const uint32_t* lo, const uint32_t* hi)
{
    SahBox b;
    for (int k = 0; k < 3; k++) {
        b.lo[k] = sah_unkey(lo[k]);
        b.hi[k] = sah_unkey(hi[k]);
    }
    return b;
}

struct DeviceBins {
    const NodeBins* b;
    __device__ SahBox box(int axis, int bin) const { return box_from_keys(b->lo[axis][bin], b->hi[axis][bin]); }
    __device__ uint32_t count(int axis, int bin) const { return b->count[axis][bin]; }
};

// Warp-aggregated accumulation: lanes that target the same record (same `key`; kNone = no contribution) are grouped with
// __match_any_sync, their boxes reduced with redux.sync (__reduce_min_sync / __reduce_max_sync on the integer keys) and only the
// group's first lane issues the atomics.  Near the root, where a whole warp (consecutive positions) belongs to one node, this
// turns 32 same-address atomics into one; deep in the tree it degenerates to one atomic per lane.  Call with all 32 lanes.
struct Group {
    unsigned peers;
    bool active, leader;
};
__device__ __forceinline__ Group group_by(uint32_t key)
{
    const unsigned lane = threadIdx.x & 31;
    Group g;
    g.peers = __match_any_sync(0xffffffffu, key);
    g.active = key != kNone;
    g.leader = g.active && (unsigned(__ffs(int(g.peers))) - 1u == lane);
    return g;
}
__device__ __forceinline__ void group_min_max(const Group& g, uint32_t* lo, uint32_t* hi, const float* vlo, const float* vhi)
{
    if (!g.active)
        return;
    for (int k = 0; k < 3; k++) {
        const uint32_t m = __reduce_min_sync(g.peers, sah_key(vlo[k]));
        const uint32_t M = __reduce_max_sync(g.peers, sah_key(vhi[k]));
        if (g.leader) {
            atomicMin(lo + k, m);
            atomicMax(hi + k, M);
        }
    }
}

__global__ void prim_box_kernel(const cge_vertex* __restrict__ verts, const uint32_t* __restrict__ tris,
    const uint32_t* __restrict__ triVertexOffset, uint32_t nTris, PrimBox* __restrict__ out, uint32_t* __restrict__ order,
    uint32_t* __restrict__ posNode, LevelNode* root)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < nTris;
    PrimBox p {};
    if (valid) {
        const uint32_t vo = triVertexOffset[i];
        const cge_vertex& a = verts[vo + tris[3 * size_t(i)]];
        const cge_vertex& b = verts[vo + tris[3 * size_t(i) + 1]];
        const cge_vertex& c = verts[vo + tris[3 * size_t(i) + 2]];
        sah_triangle_bounds(a.position, b.position, c.position, p.lo, p.hi, p.c);
        out[i] = p;
        order[i] = i;
        posNode[i] = 0;
    }
    const Group g = group_by(valid ? 0u : kNone);
    group_min_max(g, root->lo, root->hi, p.lo, p.hi);
    group_min_max(g, root->clo, root->chi, p.c, p.c);
}

__global__ void init_nodes_kernel(LevelNode* nodes, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    LevelNode& nd = nodes[i];
    nd.beg = nd.end = 0;
    for (int k = 0; k < 3; k++) {
        nd.lo[k] = nd.clo[k] = kKeyEmptyLo;
        nd.hi[k] = nd.chi[k] = kKeyEmptyHi;
    }
    nd.parent = kNone;
    nd.side = 0;
}

__global__ void init_bins_kernel(NodeBins* bins, uint32_t nNodes)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; // one thread per (node, axis, bin)
    if (i >= nNodes * 3u * kSahBins)
        return;
    const uint32_t node = i / (3u * kSahBins), r = i % (3u * kSahBins), axis = r / kSahBins, bin = r % kSahBins;
    NodeBins& nb = bins[node];
    for (int k = 0; k < 3; k++) {
        nb.lo[axis][bin][k] = kKeyEmptyLo;
        nb.hi[axis][bin][k] = kKeyEmptyHi;
    }
    nb.count[axis][bin] = 0;
}

__global__ void bin_kernel(const PrimBox* __restrict__ prims, const uint32_t* __restrict__ order, const uint32_t* __restrict__ posNode,
    const LevelNode* __restrict__ nodes, NodeBins* bins, uint32_t n, uint32_t depth)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a = i < n ? posNode[i] : kNone;
    if (a != kNone && (nodes[a].end - nodes[a].beg <= 1 || depth >= kSahMaxDepth))
        a = kNone; // single primitive or depth cap: no SAH split is evaluated
    PrimBox p {};
    if (a != kNone)
        p = prims[order[i]];
    for (int axis = 0; axis < 3; axis++) {
        uint32_t key = kNone;
        int b = 0;
        if (a != kNone) {
            const float lo = sah_unkey(nodes[a].clo[axis]);
            const float ext = sah_unkey(nodes[a].chi[axis]) - lo;
            if (ext > 0.0f) {
                b = sah_bin_of(p.c[axis], lo, sah_bin_scale(ext));
                key = a * kSahBins + uint32_t(b);
            }
        }
        const Group g = group_by(key);
        if (!g.active)
            continue;
        NodeBins& nb = bins[a];
        group_min_max(g, nb.lo[axis][b], nb.hi[axis][b], p.lo, p.hi);
        if (g.leader)
            atomicAdd(&nb.count[axis][b], unsigned(__popc(g.peers)));
    }
}

__global__ void decide_kernel(const LevelNode* __restrict__ nodes, const NodeBins* __restrict__ bins, Decision* __restrict__ dec,
    uint32_t* __restrict__ splits, uint32_t nNodes, uint32_t depth)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nNodes)
        return;
    const LevelNode& nd = nodes[a];
    const SahDecision d = sah_decide(nd.end - nd.beg, depth, box_from_keys(nd.lo, nd.hi), box_from_keys(nd.clo, nd.chi), DeviceBins { bins + a });
    dec[a] = Decision { d.kind, d.axis, d.bin, d.lo, d.scale };
    splits[a] = d.kind == kSahLeaf ? 0u : 1u;
}

// "goes left" flag of every position: 1 for primitives that stay in the left part of their node's range (all primitives
// of a node that becomes a leaf, and of ranges that are already leaves, so that they keep their place)
__global__ void flag_kernel(const PrimBox* __restrict__ prims, const uint32_t* __restrict__ order, const uint32_t* __restrict__ posNode,
    const LevelNode* __restrict__ nodes, const Decision* __restrict__ dec, uint32_t* __restrict__ flags, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const uint32_t a = posNode[i];
    uint32_t f = 1;
    if (a != kNone) {
        const Decision d = dec[a];
        if (d.kind == kSahSplitBin)
            f = sah_bin_of(prims[order[i]].c[d.axis], d.lo, d.scale) <= d.bin ? 1u : 0u;
        else if (d.kind == kSahSplitMiddle)
            f = i < nodes[a].beg + (nodes[a].end - nodes[a].beg) / 2 ? 1u : 0u;
    }
    flags[i] = f;
}

// one thread per node of the level: record the decision in the parent, open the two children in the next level
__global__ void emit_kernel(const LevelNode* __restrict__ nodes, const Decision* __restrict__ dec, const uint32_t* __restrict__ splitScan,
    const uint32_t* __restrict__ flagScan, LevelNode* __restrict__ next, BfsNode* __restrict__ bfs, uint32_t bfsBase, uint32_t nNodes,
    uint32_t* __restrict__ rootRef, uint32_t* __restrict__ nLeaves)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nNodes)
        return;
    const LevelNode& nd = nodes[a];
    const Decision d = dec[a];
    uint32_t ref;
    if (d.kind == kSahLeaf) {
        ref = fast_leaf_ref(nd.beg, nd.end - nd.beg);
        atomicAdd(nLeaves, 1u);
    } else {
        const uint32_t me = bfsBase + splitScan[a]; // breadth-first index of this inner node
        ref = me;
        const uint32_t left = flagScan[nd.end] - flagScan[nd.beg]; // flagScan has nPrims + 1 entries
        const uint32_t mid = nd.beg + left;
        LevelNode& l = next[2 * splitScan[a]];
        LevelNode& r = next[2 * splitScan[a] + 1];
        l.beg = nd.beg, l.end = mid, l.parent = me, l.side = 0;
        r.beg = mid, r.end = nd.end, r.parent = me, r.side = 1;
    }
    if (nd.parent == kNone) {
        *rootRef = ref;
    } else {
        BfsNode& p = bfs[nd.parent];
        p.child[nd.side] = ref;
        for (int k = 0; k < 3; k++) {
            p.box[nd.side][k] = sah_unkey(nd.lo[k]);
            p.box[nd.side][3 + k] = sah_unkey(nd.hi[k]);
        }
    }
}

__global__ void partition_kernel(const PrimBox* __restrict__ prims, const uint32_t* __restrict__ order, const uint32_t* __restrict__ posNode,
    const LevelNode* __restrict__ nodes, const Decision* __restrict__ dec, const uint32_t* __restrict__ splitScan,
    const uint32_t* __restrict__ flags, const uint32_t* __restrict__ flagScan, uint32_t* __restrict__ order2, uint32_t* __restrict__ posNode2,
    LevelNode* next, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t child = kNone;
    PrimBox p {};
    if (i < n) {
        const uint32_t a = posNode[i];
        const uint32_t prim = order[i];
        if (a == kNone || dec[a].kind == kSahLeaf) { // already in a leaf, or its node just became one: stays where it is
            order2[i] = prim;
            posNode2[i] = kNone;
        } else {
            const LevelNode& nd = nodes[a];
            const uint32_t leftBefore = flagScan[i] - flagScan[nd.beg]; // primitives of this node before i that go left
            const uint32_t leftTotal = flagScan[nd.end] - flagScan[nd.beg];
            const bool goesLeft = flags[i] != 0;
            const uint32_t pos = goesLeft ? nd.beg + leftBefore : nd.beg + leftTotal + ((i - nd.beg) - leftBefore);
            child = 2 * splitScan[a] + (goesLeft ? 0u : 1u);
            order2[pos] = prim;
            posNode2[pos] = child;
            p = prims[prim];
        }
    }
    const Group g = group_by(child);
    if (!g.active)
        return;
    LevelNode& c = next[child];
    group_min_max(g, c.lo, c.hi, p.lo, p.hi);
    group_min_max(g, c.clo, c.chi, p.c, p.c);
}

// ---- exclusive prefix sum over n + 1 outputs (out[n] = total), 1024 items per block ------------------------------------------
constexpr uint32_t kScanBlock = 256, kScanItems = 4, kScanTile = kScanBlock * kScanItems;

__global__ void scan_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ tileSums, uint32_t n)
{
    __shared__ uint32_t warpSums[kScanBlock / 32];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems], sum = 0;
    for (uint32_t k = 0; k < kScanItems; k++) {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    uint32_t incl = sum;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= uint32_t(off))
            incl += t;
    }
    if (lane == 31)
        warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScanBlock / 32 ? warpSums[lane] : 0u;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= uint32_t(off))
                w += t;
        }
        if (lane < kScanBlock / 32)
            warpSums[lane] = w; // inclusive over warps
    }
    __syncthreads();
    uint32_t excl = incl - sum + (warp ? warpSums[warp - 1] : 0u);
    for (uint32_t k = 0; k < kScanItems; k++) {
        if (base + k < n)
            out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == kScanBlock - 1)
        tileSums[blockIdx.x] = excl;
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ tileOffsets, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] += tileOffsets[i / kScanTile];
    if (i == 0)
        out[n] = tileOffsets[(n + kScanTile - 1) / kScanTile]; // total
}

struct Scan {
    // level 0 tile sums, level 1 tile sums, ... (each level 1024x smaller), allocated once for the largest input
    std::vector<uint32_t*> sums, offsets;
    cudaError_t init(uint32_t maxN)
    {
        uint32_t n = maxN;
        while (true) {
            const uint32_t tiles = (n + kScanTile - 1) / kScanTile;
            uint32_t *s = nullptr, *o = nullptr;
            cudaError_t e = cudaMalloc(&s, (size_t(tiles) + 1) * sizeof(uint32_t));
            if (e == cudaSuccess)
                e = cudaMalloc(&o, (size_t(tiles) + 2) * sizeof(uint32_t));
            if (e != cudaSuccess)
                return e;
            sums.push_back(s);
            offsets.push_back(o);
            if (tiles <= 1)
                break;
            n = tiles;
        }
        return cudaSuccess;
    }
    ~Scan()
    {
        for (auto* p : sums)
            cudaFree(p);
        for (auto* p : offsets)
            cudaFree(p);
    }
    // out has n + 1 entries
    void run(const uint32_t* in, uint32_t* out, uint32_t n, cudaStream_t st, size_t level = 0)
    {
        const uint32_t tiles = (n + kScanTile - 1) / kScanTile;
        scan_tile_kernel<<<std::max(tiles, 1u), kScanBlock, 0, st>>>(in, out, sums[level], n);
        if (tiles <= 1) {
            // total = the single tile's sum
            cudaMemcpyAsync(out + n, sums[level], sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
            return;
        }
        run(sums[level], offsets[level], tiles, st, level + 1); // offsets[level][tiles] = total
        scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, offsets[level], n);
    }
};

// ---- depth-first renumbering -----------------------------------------------------------------------------------------------
__global__ void subtree_size_kernel(const BfsNode* __restrict__ bfs, uint32_t* __restrict__ size, uint32_t beg, uint32_t end)
{
    const uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end)
        return;
    uint32_t s = 1;
    for (int side = 0; side < 2; side++)
        if (!(bfs[i].child[side] & kFastLeafBit))
            s += size[bfs[i].child[side]];
    size[i] = s;
}
__global__ void preorder_index_kernel(const BfsNode* __restrict__ bfs, const uint32_t* __restrict__ size, uint32_t* __restrict__ dfs,
    uint32_t beg, uint32_t end)
{
    const uint32_t i = beg + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end)
        return;
    const uint32_t l = bfs[i].child[0], r = bfs[i].child[1];
    uint32_t next = dfs[i] + 1;
    if (!(l & kFastLeafBit)) {
        dfs[l] = next;
        next += size[l];
    }
    if (!(r & kFastLeafBit))
        dfs[r] = next;
}
__global__ void emit_dfs_kernel(const BfsNode* __restrict__ bfs, const uint32_t* __restrict__ dfs, FastNode* __restrict__ out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const BfsNode b = bfs[i];
    FastNode f;
    for (int k = 0; k < 3; k++) {
        f.l_lo[k] = b.box[0][k], f.l_hi[k] = b.box[0][3 + k];
        f.r_lo[k] = b.box[1][k], f.r_hi[k] = b.box[1][3 + k];
    }
    f.left = (b.child[0] & kFastLeafBit) ? b.child[0] : dfs[b.child[0]];
    f.right = (b.child[1] & kFastLeafBit) ? b.child[1] : dfs[b.child[1]];
    out[dfs[i]] = f;
}

template <typename T>
struct Dev {
    T* p = nullptr;
    ~Dev() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T)); }
};

} // namespace

#define SAH_TRY(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            if (err)                                                                           \
                *err = std::string("GPU SAH build: ") + cudaGetErrorString(e_);                \
            return false;                                                                      \
        }                                                                                      \
    } while (0)

bool sah_gpu_supported(const cge_scene_desc& d)
{
    if (d.n_spheres || d.n_triangles == 0 || d.n_triangles >= (1u << 28))
        return false;
    // atomic min / max on keys order NaN outside the number line: scenes with non-finite coordinates take the host builder
    for (uint32_t m = 0; m < d.n_meshes; m++) {
        const cge_mesh_desc& md = d.meshes[m];
        for (uint32_t v = 0; v < md.vertex_count; v++)
            for (int k = 0; k < 3; k++)
                if (!std::isfinite(d.vertices[md.vertex_offset + v].position[k]))
                    return false;
    }
    return true;
}

bool build_sah_bvh_gpu(const cge_scene_desc& d, FastBvh& out, float* buildMs, std::string* err)
{
    out = FastBvh {};
    const uint32_t n = d.n_triangles;
    cudaStream_t st = nullptr; // legacy default stream: scene creation is synchronous anyway
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    SAH_TRY(cudaEventCreate(&ev0));
    SAH_TRY(cudaEventCreate(&ev1));
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard()
        {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } evGuard { ev0, ev1 };

    // CGE_TIMING=1: host wall clock of every phase on stderr (development aid)
    const bool trace = std::getenv("CGE_TIMING") != nullptr;
    auto wall0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace)
            return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[sah-gpu] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - wall0).count());
        wall0 = now;
    };

    // inputs
    std::vector<uint32_t> triVertexOffset(n);
    for (uint32_t m = 0; m < d.n_meshes; m++)
        for (uint32_t t = 0; t < d.meshes[m].triangle_count; t++)
            triVertexOffset[d.meshes[m].triangle_offset + t] = d.meshes[m].vertex_offset;
    Dev<cge_vertex> dVerts;
    Dev<uint32_t> dTris, dVo, order, order2, posNode, posNode2, flags, flagScan, splits, splitScan, counters, subtree, dfs;
    Dev<PrimBox> prims;
    Dev<LevelNode> levelA, levelB;
    Dev<NodeBins> bins;
    Dev<Decision> dec;
    Dev<BfsNode> bfs;
    Dev<FastNode> outNodes;
    SAH_TRY(dVerts.alloc(d.n_vertices));
    SAH_TRY(dTris.alloc(size_t(n) * 3));
    SAH_TRY(dVo.alloc(n));
    SAH_TRY(cudaMemcpyAsync(dVerts.p, d.vertices, size_t(d.n_vertices) * sizeof(cge_vertex), cudaMemcpyHostToDevice, st));
    SAH_TRY(cudaMemcpyAsync(dTris.p, d.triangles, size_t(n) * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SAH_TRY(cudaMemcpyAsync(dVo.p, triVertexOffset.data(), size_t(n) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SAH_TRY(prims.alloc(n));
    for (auto* b : { &order, &order2, &posNode, &posNode2, &flags })
        SAH_TRY(b->alloc(n));
    SAH_TRY(flagScan.alloc(size_t(n) + 1));
    // a level has at most n nodes (every node owns at least one primitive); inner nodes in total: at most n - 1
    const uint32_t maxLevelNodes = n + 1;
    SAH_TRY(levelA.alloc(maxLevelNodes));
    SAH_TRY(levelB.alloc(maxLevelNodes));
    SAH_TRY(dec.alloc(maxLevelNodes));
    SAH_TRY(splits.alloc(maxLevelNodes));
    SAH_TRY(splitScan.alloc(size_t(maxLevelNodes) + 1));
    SAH_TRY(bfs.alloc(n));
    SAH_TRY(counters.alloc(4)); // [0] root reference, [1] leaves
    SAH_TRY(cudaMemsetAsync(counters.p, 0, 4 * sizeof(uint32_t), st));
    Scan scan;
    SAH_TRY(scan.init(maxLevelNodes));
    // bins are only touched for nodes with >= 2 primitives, of which a level has at most n / 2; nodes of a level are numbered
    // in creation order, so the buffer is sized for the widest level that can occur
    size_t binsCap = 0;
    lap("upload + allocations");

    SAH_TRY(cudaEventRecord(ev0, st));
    init_nodes_kernel<<<1, 32, 0, st>>>(levelA.p, 1);
    prim_box_kernel<<<(n + 255) / 256, 256, 0, st>>>(dVerts.p, dTris.p, dVo.p, n, prims.p, order.p, posNode.p, levelA.p);
    {
        const uint32_t range[2] = { 0u, n };
        SAH_TRY(cudaMemcpyAsync(levelA.p, range, sizeof(range), cudaMemcpyHostToDevice, st)); // root: beg = 0, end = n
    }

    lap("primitive boxes + root");
    std::vector<uint32_t> levelStart { 0 }; // breadth-first index of the first inner node of every level
    uint32_t nLevelNodes = 1, nInner = 0, depth = 0;
    LevelNode *cur = levelA.p, *next = levelB.p;
    uint32_t *ord = order.p, *ord2 = order2.p, *pn = posNode.p, *pn2 = posNode2.p;
    while (nLevelNodes > 0) {
        if (depth > 4 * kSahMaxDepth) {
            if (err)
                *err = "GPU SAH build: runaway depth";
            return false;
        }
        if (nLevelNodes > binsCap) {
            cudaFree(bins.p);
            bins.p = nullptr;
            binsCap = std::max<size_t>(nLevelNodes, binsCap * 2);
            binsCap = std::min<size_t>(binsCap, maxLevelNodes);
            SAH_TRY(bins.alloc(binsCap));
        }
        init_bins_kernel<<<(nLevelNodes * 3u * kSahBins + 255) / 256, 256, 0, st>>>(bins.p, nLevelNodes);
        bin_kernel<<<(n + 255) / 256, 256, 0, st>>>(prims.p, ord, pn, cur, bins.p, n, depth);
        decide_kernel<<<(nLevelNodes + 127) / 128, 128, 0, st>>>(cur, bins.p, dec.p, splits.p, nLevelNodes, depth);
        scan.run(splits.p, splitScan.p, nLevelNodes, st);
        uint32_t nSplit = 0;
        SAH_TRY(cudaMemcpyAsync(&nSplit, splitScan.p + nLevelNodes, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        flag_kernel<<<(n + 255) / 256, 256, 0, st>>>(prims.p, ord, pn, cur, dec.p, flags.p, n);
        scan.run(flags.p, flagScan.p, n, st);
        SAH_TRY(cudaStreamSynchronize(st)); // nSplit sizes the next level
        if (nInner + nSplit > n) {
            if (err)
                *err = "GPU SAH build: node count overflow";
            return false;
        }
        if (nSplit)
            init_nodes_kernel<<<(2 * nSplit + 255) / 256, 256, 0, st>>>(next, 2 * nSplit);
        emit_kernel<<<(nLevelNodes + 127) / 128, 128, 0, st>>>(cur, dec.p, splitScan.p, flagScan.p, next, bfs.p, nInner, nLevelNodes,
            counters.p, counters.p + 1);
        partition_kernel<<<(n + 255) / 256, 256, 0, st>>>(prims.p, ord, pn, cur, dec.p, splitScan.p, flags.p, flagScan.p, ord2, pn2, next, n);
        std::swap(ord, ord2);
        std::swap(pn, pn2);
        std::swap(cur, next);
        nInner += nSplit;
        levelStart.push_back(nInner);
        if (trace) {
            char label[64];
            std::snprintf(label, sizeof(label), "level %u: %u nodes, %u split", depth, nLevelNodes, nSplit);
            lap(label);
        }
        nLevelNodes = 2 * nSplit;
        depth++;
    }
    SAH_TRY(cudaGetLastError());

    // depth-first renumbering (levels bottom-up for the subtree sizes, top-down for the pre-order index)
    if (nInner) {
        SAH_TRY(subtree.alloc(nInner));
        SAH_TRY(dfs.alloc(nInner));
        SAH_TRY(outNodes.alloc(nInner));
        SAH_TRY(cudaMemsetAsync(dfs.p, 0, sizeof(uint32_t), st)); // the root (breadth-first index 0) keeps index 0
        const int nLevels = int(levelStart.size()) - 1;
        for (int l = nLevels - 1; l >= 0; l--) {
            const uint32_t beg = levelStart[l], end = levelStart[l + 1];
            if (end > beg)
                subtree_size_kernel<<<(end - beg + 255) / 256, 256, 0, st>>>(bfs.p, subtree.p, beg, end);
        }
        for (int l = 0; l < nLevels; l++) {
            const uint32_t beg = levelStart[l], end = levelStart[l + 1];
            if (end > beg)
                preorder_index_kernel<<<(end - beg + 255) / 256, 256, 0, st>>>(bfs.p, subtree.p, dfs.p, beg, end);
        }
        emit_dfs_kernel<<<(nInner + 255) / 256, 256, 0, st>>>(bfs.p, dfs.p, outNodes.p, nInner);
    }
    SAH_TRY(cudaEventRecord(ev1, st));
    lap("depth-first renumbering");

    out.nodes.resize(nInner);
    out.prim_order.resize(n);
    uint32_t cnt[4] = {};
    if (nInner)
        SAH_TRY(cudaMemcpyAsync(out.nodes.data(), outNodes.p, size_t(nInner) * sizeof(FastNode), cudaMemcpyDeviceToHost, st));
    SAH_TRY(cudaMemcpyAsync(out.prim_order.data(), ord, size_t(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SAH_TRY(cudaMemcpyAsync(cnt, counters.p, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    SAH_TRY(cudaStreamSynchronize(st));
    SAH_TRY(cudaGetLastError());
    lap("download");
    out.root = (cnt[0] & kFastLeafBit) ? cnt[0] : 0u; // the root inner node has pre-order index 0
    out.n_leaves = cnt[1];
    out.depth = depth;
    if (buildMs)
        SAH_TRY(cudaEventElapsedTime(buildMs, ev0, ev1));
    return true;
}

} // namespace cge

// rz_bvh_wide.cu — BVH2 -> BVH4: collapses the binary tree of either builder (host binned SAH, rz_context.cu; device LBVH,
// rz_bvh_build.cu) into the 4-wide nodes the traversal kernel K3 walks (RzBvh4Node, rz_device.cuh).
//
// Every internal node at EVEN depth becomes a wide node whose children are its grandchildren (a child that is a leaf stays a
// child).  Nodes at odd depth are absorbed.  Depth parity comes from a parent array and a walk to the root (depth <= 96 in
// the LBVH, ~20 in the SAH tree), so the collapse is three small kernels on the caller's stream and needs no host round
// trip: the device-built tree stays device-built.  Wide node i lives at the index of the binary node it was made from, so
// child references need no translation (the slots of absorbed nodes stay empty: 128 B x n, 12.8 MB at 100k spheres).
// The closest hit does not depend on the tree, so images are unchanged bit for bit.
#include "rz_device.cuh"

namespace {

__global__ void __launch_bounds__(256) wide_parents(const RzBvhNode *n2, uint32_t n_nodes, int *parent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    if (i == 0) parent[0] = -1;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int ch = n2[i].child[c];
        if (ch >= 0 && (uint32_t)ch < n_nodes && n2[i].cnt[c] == 0u) parent[ch] = (int)i;
    }
}

__global__ void __launch_bounds__(256) wide_collapse(const RzBvhNode *n2, uint32_t n_nodes, const int *parent, RzBvh4Node *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    RzBvh4Node w;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        w.lox[c] = w.loy[c] = w.loz[c] = INFINITY;
        w.hix[c] = w.hiy[c] = w.hiz[c] = -INFINITY;
        w.child[c] = ~0; w.cnt[c] = 0u;
    }
    // depth parity: nodes that were never referenced (inside a collapsed LBVH subtree) have parent == -2 and stay empty
    int odd = 0, j = (int)i, steps = 0;
    bool reachable = true;
    while (j != 0) {
        const int p = parent[j];
        if (p < 0 || ++steps > 4096) { reachable = false; break; }
        j = p; odd ^= 1;
    }
    if (reachable && !odd) {
        int k = 0;
        auto put = [&](const RzBvhNode &m, int c) {
            if (m.child[c] < 0 && m.cnt[c] == 0u) return;            // unused slot of the binary node
            w.lox[k] = m.lox[c]; w.hix[k] = m.hix[c]; w.loy[k] = m.loy[c]; w.hiy[k] = m.hiy[c]; w.loz[k] = m.loz[c]; w.hiz[k] = m.hiz[c];
            w.child[k] = m.child[c]; w.cnt[k] = m.cnt[c];
            k++;
        };
        const RzBvhNode me = n2[i];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (me.child[c] >= 0 && me.cnt[c] == 0u) {                // internal child: absorbed, its children move up
                const RzBvhNode m = n2[me.child[c]];
                put(m, 0);
                put(m, 1);
            } else {
                put(me, c);                                           // leaf (or unused) child
            }
        }
    }
    out[i] = w;
}

__global__ void wide_fill(int *p, uint32_t n, int v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

extern "C" size_t rz_bvh_wide_scratch_bytes(uint32_t n_nodes) { return (size_t)(n_nodes ? n_nodes : 1) * sizeof(int); }

// n2: binary nodes [n_nodes], root 0 (both child boxes in the parent; child < 0 && cnt > 0: leaf).  out: [n_nodes] wide nodes.
extern "C" cudaError_t rz_bvh_wide_collapse(const RzBvhNode *n2, uint32_t n_nodes, void *scratch, RzBvh4Node *out, cudaStream_t stream) {
    if (n_nodes == 0) return cudaErrorInvalidValue;
    int *parent = static_cast<int *>(scratch);
    const unsigned grid = (n_nodes + 255u) / 256u;
    wide_fill<<<grid, 256, 0, stream>>>(parent, n_nodes, -2);
    wide_parents<<<grid, 256, 0, stream>>>(n2, n_nodes, parent);
    wide_collapse<<<grid, 256, 0, stream>>>(n2, n_nodes, parent, out);
    return cudaGetLastError();
}

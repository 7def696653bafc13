// Test-only: the product's host-side tree builders (rayz_b200/csrc/rz_host_bvh.hpp) compiled on their own, behind a tiny C API
// for tests/test_host_bvh_cpu.py.  Nothing here runs on a GPU or renders anything.
#include "../../rayz_b200/csrc/rz_host_bvh.hpp"

static RzScene scene_of(uint32_t n, const double *c, const double *v, const double *r) {
    RzScene sc;
    memset(&sc, 0, sizeof sc);
    sc.n_spheres = n;
    sc.sphere_center = c; sc.sphere_velocity = v; sc.sphere_radius = r;
    return sc;
}

// K3's binned-SAH tree as rayz_cuda_upload_scene builds it; returns the node count (or -1 if `nodes_cap` is too small)
extern "C" int hostbvh_sah(uint32_t n, const double *c, const double *v, const double *r, int leaf, double node_cost, RzBvhNode *nodes_out,
                           uint32_t nodes_cap, uint32_t *order_out) {
    const RzScene sc = scene_of(n, c, v, r);
    SahBuilder sb;
    sb.LEAF = leaf;
    sb.node_cost = node_cost;
    sb.p.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        sb.p[i].b = sphere_box(sc, i);
        for (int a = 0; a < 3; a++) sb.p[i].c[a] = 0.5 * (sb.p[i].b.lo[a] + sb.p[i].b.hi[a]);
        sb.p[i].s = i;
    }
    sb.run();
    if (sb.nodes.size() > nodes_cap) return -1;
    memcpy(nodes_out, sb.nodes.data(), sb.nodes.size() * sizeof(RzBvhNode));
    memcpy(order_out, sb.order.data(), sb.order.size() * sizeof(uint32_t));
    return (int)sb.nodes.size();
}

// K0's reference-shaped tree (BVH.build, hit.zig:130-161); returns the node count (or -1)
extern "C" int hostbvh_ref(uint32_t n, const double *c, const double *v, const double *r, RzRefNode *nodes_out, uint32_t nodes_cap,
                           uint32_t *order_out) {
    const RzScene sc = scene_of(n, c, v, r);
    RefBuilder rb;
    rb.h.resize(n);
    for (uint32_t i = 0; i < n; i++) { rb.h[i].b = sphere_box(sc, i); rb.h[i].s = i; }
    if (n) rb.build(0, n);
    if (rb.nodes.size() > nodes_cap) return -1;
    memcpy(nodes_out, rb.nodes.data(), rb.nodes.size() * sizeof(RzRefNode));
    for (uint32_t i = 0; i < n; i++) order_out[i] = rb.h[i].s;
    return (int)rb.nodes.size();
}

extern "C" void hostbvh_sphere_boxes(uint32_t n, const double *c, const double *v, const double *r, double *lo, double *hi) {
    const RzScene sc = scene_of(n, c, v, r);
    for (uint32_t i = 0; i < n; i++) {
        const Box b = sphere_box(sc, i);
        for (int a = 0; a < 3; a++) { lo[3 * i + a] = b.lo[a]; hi[3 * i + a] = b.hi[a]; }
    }
}

extern "C" void hostbvh_next(const float *in, uint32_t n, float *down, float *up) {
    for (uint32_t i = 0; i < n; i++) { down[i] = SahBuilder::next_down(in[i]); up[i] = SahBuilder::next_up(in[i]); }
}

extern "C" void hostbvh_round(const double *in, uint32_t n, float *down, float *up) {
    for (uint32_t i = 0; i < n; i++) { down[i] = SahBuilder::down(in[i]); up[i] = SahBuilder::up(in[i]); }
}

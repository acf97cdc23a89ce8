// Contract check of the traversal records on the CPU (no device is touched): the encoder (bvh.cuh: trav_encode_inner,
// trav_encode_quads -- what flatten_emit_kernel / pack_quads_kernel store) and the decoder (trav_axis_planes2,
// trav_axis_vals4 -- what trace_persistent_kernel evaluates) are the very functions the kernels inline, compiled here
// for the host.  For random nodes, child boxes and rays it asserts the obligation the default traversal rests on:
//
//     the reference's slab test (raytracer.glsl:182-237) passes on the EXACT child box
//         ==>  the compressed test passes, and its entry distance is not larger than the reference's
//
// for the pair step and for the four-slot wide step, with and without a pruning limit.  Built and run by
// tests/test_trav_records_cpu.py (nvcc -Xcompiler -ffp-contract=off,-frounding-math; host code only).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>

#include "../../realtimeraytracing_b200/csrc/bvh.cuh"

struct Box { float lo[3], hi[3]; };

// the reference's test, as trace.cu: intersect_box restates it (every op one rounded fp32 op)
static bool ref_slab(const float o[3], const float inv[3], const Box& b, float* t_entry) {
    float tMin = -INFINITY, tMax = INFINITY;
    for (int k = 0; k < 3; ++k) {
        const float t1 = trav_mul(trav_sub(b.lo[k], o[k]), inv[k]), t2 = trav_mul(trav_sub(b.hi[k], o[k]), inv[k]);
        if (k == 0) { tMin = fminf(t1, t2); tMax = fmaxf(t1, t2); }
        else { tMin = fmaxf(tMin, fminf(t1, t2)); tMax = fminf(tMax, fmaxf(t1, t2)); }
        if (k < 2 && (tMax < 0.f || tMin > tMax)) return false;
    }
    *t_entry = tMin;
    return tMax >= 0.f && tMin <= tMax;
}

static float4 lo4(const Box& b) { return make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]); }
static float2 hi2(const Box& b) { return make_float2(b.hi[1], b.hi[2]); }

// --shrink: move every encoded min plane two steps up and every max plane two steps down -- boxes that no longer contain
// their children; the check must notice (it is how the test knows that the check can fail)
static uint32_t shrink(uint32_t quad) {
    uint32_t out = 0;
    for (int i = 0; i < 4; ++i) {
        int b = (int)((quad >> (8 * i)) & 0xFFu);
        b = (i & 1) ? (b - 2 < 0 ? 0 : b - 2) : (b + 2 > 255 ? 255 : b + 2);
        out |= (uint32_t)b << (8 * i);
    }
    return out;
}

int main(int argc, char** argv) {
    const long cases = argc > 1 ? std::atol(argv[1]) : 400000;
    const bool broken = argc > 2 && std::string(argv[2]) == "--shrink";
    const bool wide = argc > 2 && std::string(argv[2]) == "--wide";  // scene scales 1e-25 ... 1e14: some nodes are unencodable
    std::mt19937_64 rng(20240611);
    auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
    auto logu = [&](double a, double b) { return std::exp(uni(std::log(a), std::log(b))); };
    long checked = 0, ref_hits = 0, bad = 0, unusable = 0;
    for (long c = 0; c < cases; ++c) {
        // a node somewhere, of some size; four boxes inside it (some flat, some touching its faces, some the node itself)
        Box node;
        const double scale = wide ? logu(1e-25, 1e14) : logu(1e-3, 1e4);
        for (int k = 0; k < 3; ++k) {
            const double centre = uni(-1.0, 1.0) * scale * (c % 3 == 0 ? 100.0 : 1.0);
            const double ext = scale * logu(1e-4, 1.0);
            node.lo[k] = (float)(centre - ext); node.hi[k] = (float)(centre + ext);
            if (!(node.lo[k] < node.hi[k])) node.hi[k] = std::nextafter(node.lo[k], INFINITY);
        }
        Box sub[4];
        for (int s = 0; s < 4; ++s)
            for (int k = 0; k < 3; ++k) {
                double a = uni(0, 1), b = uni(0, 1);
                if (a > b) std::swap(a, b);
                const int kind = (int)(rng() % 8);
                if (kind == 0) a = 0;
                if (kind == 1) b = 1;
                if (kind == 2) b = a;  // flat
                if (kind == 3) { a = 0; b = 1; }
                const double w = (double)node.hi[k] - (double)node.lo[k];
                float lo = (float)((double)node.lo[k] + a * w), hi = (float)((double)node.lo[k] + b * w);
                lo = fminf(fmaxf(lo, node.lo[k]), node.hi[k]); hi = fminf(fmaxf(hi, lo), node.hi[k]);
                sub[s].lo[k] = lo; sub[s].hi[k] = hi;
            }
        // a ray from near or far, some directions almost parallel to an axis (never exactly: those rays take the exact path)
        float o[3], d[3], inv[3];
        const double far = (c % 5 == 0) ? 1e3 : 3.0;
        for (int k = 0; k < 3; ++k) {
            o[k] = (float)(0.5 * ((double)node.lo[k] + node.hi[k]) + uni(-1, 1) * far * scale);
            double dk = uni(-1, 1);
            if (rng() % 6 == 0) dk *= logu(1e-9, 1e-2);
            if (dk == 0.0) dk = 1e-3;
            d[k] = (float)dk;
        }
        if (c % 2 == 0) {  // aim at the node, so that hits are common
            for (int k = 0; k < 3; ++k) {
                const double target = (double)node.lo[k] + uni(0, 1) * ((double)node.hi[k] - node.lo[k]);
                const double dk = target - o[k];
                d[k] = (float)(dk == 0.0 ? 1e-3 : dk);
            }
        }
        bool fast = true;
        for (int k = 0; k < 3; ++k) { inv[k] = 1.0f / d[k]; fast = fast && fabsf(inv[k]) < 1.8e19f; }
        if (!fast) continue;  // trace.cu sends such rays to the reference's own slab test on the decoded boxes
        const float limit = (c % 3 == 1) ? (float)(logu(1e-3, 1e3) * scale) : INFINITY;

        // ---- pair step: children = sub[0], sub[1] ----
        uint4 o0, o1;
        trav_encode_inner(lo4(node), hi2(node), lo4(sub[0]), hi2(sub[0]), lo4(sub[1]), hi2(sub[1]), false, false, 7u, o0, o1);
        const uint32_t flags = o0.w >> 24;
        if (flags & 4u) { ++unusable; continue; }
        if (broken) { o1.x = shrink(o1.x); o1.y = shrink(o1.y); o1.z = shrink(o1.z); }
        float tn[2] = {-INFINITY, -INFINITY}, tf[2] = {INFINITY, INFINITY};
        trav_axis_planes2(trav_u2f(o0.x), o0.w & 0xFFu, o1.x, o[0], inv[0], tn[0], tf[0], tn[1], tf[1]);
        trav_axis_planes2(trav_u2f(o0.y), (o0.w >> 8) & 0xFFu, o1.y, o[1], inv[1], tn[0], tf[0], tn[1], tf[1]);
        trav_axis_planes2(trav_u2f(o0.z), (o0.w >> 16) & 0xFFu, o1.z, o[2], inv[2], tn[0], tf[0], tn[1], tf[1]);
        for (int s = 0; s < 2; ++s) {
            float te;
            const bool ref = ref_slab(o, inv, sub[s], &te) && !(te > limit);
            const bool got = tf[s] >= 0.f && tn[s] <= tf[s] && !(tn[s] > limit);
            ++checked; ref_hits += ref;
            if (ref && (!got || tn[s] > te)) {
                if (++bad <= 5) std::fprintf(stderr, "pair step, case %ld slot %d: reference passes (t %.9g), compressed %s (t %.9g)\n", c, s, te, got ? "passes" : "FAILS", tn[s]);
            }
        }

        // ---- wide step: slots = sub[0..3] ----
        uint4 o2, o3;
        bool ok = true;
        const float4 slo[4] = {lo4(sub[0]), lo4(sub[1]), lo4(sub[2]), lo4(sub[3])};
        const float2 shi[4] = {hi2(sub[0]), hi2(sub[1]), hi2(sub[2]), hi2(sub[3])};
        trav_encode_quads(lo4(node), hi2(node), slo, shi, 1u, 3u, o2, o3, ok);
        if (!ok) { ++unusable; continue; }
        if (broken) { o2.x = shrink(o2.x); o2.y = shrink(o2.y); o2.z = shrink(o2.z); o3.x = shrink(o3.x); o3.y = shrink(o3.y); o3.z = shrink(o3.z); }
        float wn[4], wf[4], vn[4], vf[4], un[4], uf[4];
        trav_axis_vals4(trav_u2f(o0.x), o0.w & 0xFFu, o2.x, o3.x, o[0], inv[0], vn, vf);
        for (int k = 0; k < 4; ++k) { wn[k] = fmaxf(vn[k], 0.f); wf[k] = fminf(vf[k], limit); }
        trav_axis_vals4(trav_u2f(o0.y), (o0.w >> 8) & 0xFFu, o2.y, o3.y, o[1], inv[1], vn, vf);
        trav_axis_vals4(trav_u2f(o0.z), (o0.w >> 16) & 0xFFu, o2.z, o3.z, o[2], inv[2], un, uf);
        for (int k = 0; k < 4; ++k) { wn[k] = fmaxf(fmaxf(wn[k], vn[k]), un[k]); wf[k] = fminf(fminf(wf[k], vf[k]), uf[k]); }
        for (int s = 0; s < 4; ++s) {
            float te;
            const bool ref = ref_slab(o, inv, sub[s], &te) && !(te > limit);
            const bool got = wn[s] <= wf[s];
            ++checked; ref_hits += ref;
            if (ref && (!got || wn[s] > fmaxf(te, 0.f))) {
                if (++bad <= 5) std::fprintf(stderr, "wide step, case %ld slot %d: reference passes (t %.9g), compressed %s (t %.9g)\n", c, s, te, got ? "passes" : "FAILS", wn[s]);
            }
        }
    }
    std::printf("trav_records_check: %ld box tests, %ld reference hits, %ld unencodable nodes skipped, %ld violations\n", checked, ref_hits, unusable, bad);
    return bad ? 1 : (ref_hits * 20 < checked ? 3 : 0);  // 3: the generator stopped producing hits
}

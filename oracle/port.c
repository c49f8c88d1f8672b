/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the reference's hit path and sample stream, operating on
 * the flat scene description of include/rayito_b200.h.  It restates the reference's
 * SEQUENTIAL algorithms literally (scalar loops, literal Rng stepping, the recursive
 * meaning of the BVH walk) and knows nothing of the CUDA implementation; every
 * function cites the reference lines it follows (Rayito_Stage7_QT/...).
 *
 * Pinned by tests/test_oracle_port.py against oracle/_ref (the unmodified reference
 * compiled from /root/reference): hit records bit-equal on seeded ray batches, Rng /
 * CMJ / camera rays bit-equal.  Because it consumes the product's flattened scene,
 * agreement with oracle/_ref also proves the host-side flattening on the CPU alone.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march, no -ffast-math): the
 * float semantics must be the reference's.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "rayito_b200.h"

#define K_RAY_TMIN 0.0001f      /* RRay.h:23 */
#define K_RAY_TMAX 1.0e30f      /* RRay.h:28 */
#define K_MAX_STEPS 50          /* RAccel.h:379 */

typedef struct { float x, y, z; } vec3;

/* ---- work counters ----------------------------------
 * What the reference's traversal DOES for a ray batch, counted where the restated
 * algorithm does it: the roofline's algorithmic bytes are defined by these numbers
 * (bench.py), and tests/test_gpu_counters.py requires the CUDA kernels' own counters
 * to equal them on the same rays.
 *   node_pops    BVH nodes taken off a traversal stack, interior and leaf, both levels
 *                (one iteration of the loops of RAccel.h:405-468 / 493-560)
 *   tri_tests    Mesh::intersectTri / doesIntersectTri calls (RMesh.h:226-249)
 *   shape_tests  plane / sphere / rectangle tests (everything but meshes)
 *   xform_evals  Ray::transformToLocal calls (RRay.h:78-81): the set itself and every
 *                shape entered, whatever its transform holds
 *   xform_keyed  ... of which on a transform with at least one key: the reference reads
 *                that key's time, scale, rotation and translation (RMath.h:681-715)
 *   xform_pairs  ... of which on a transform with two or more keys (a key PAIR may be read)
 *   max_stack    largest number of live stack entries any ray reached (RAccel.h:379: 50) */
typedef struct
{
    uint64_t node_pops, tri_tests, shape_tests, xform_evals, xform_keyed, xform_pairs, max_stack_mesh, max_stack_top;
} port_work_t;
static port_work_t g_work;

void port_work_reset(void) { memset(&g_work, 0, sizeof(g_work)); }
void port_work_get(uint64_t* out8) { memcpy(out8, &g_work, sizeof(g_work)); }

/* ---- RMath.h:180-360 ----------------------------------------------------- */
static vec3 v3(float x, float y, float z) { vec3 r = { x, y, z }; return r; }
static vec3 vadd(vec3 a, vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static vec3 vsub(vec3 a, vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static vec3 vmul(vec3 a, vec3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static vec3 vdiv(vec3 a, vec3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
static vec3 vscale(vec3 a, float f) { return v3(f * a.x, f * a.y, f * a.z); }
static vec3 vneg(vec3 a) { return v3(-a.x, -a.y, -a.z); }
static float vdot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static vec3 vcross(vec3 a, vec3 b)
{
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static float vlen2(vec3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static float vlen(vec3 a) { return sqrtf(vlen2(a)); }
/* Vector::normalize, RMath.h:194 */
static vec3 vnormalized(vec3 a)
{
    float len = vlen(a);
    if (len > 0) { a.x /= len; a.y /= len; a.z /= len; }
    return a;
}
/* std::min / std::max */
static float fmin_std(float a, float b) { return (b < a) ? b : a; }
static float fmax_std(float a, float b) { return (a < b) ? b : a; }

/* ---- Quaternion / Transform, RMath.h:536-549, 576-580, 681-715, 814-884 ---- */
typedef struct { float w; vec3 v; } quat;

static vec3 qrot(quat q, vec3 v)
{
    vec3 t = vscale(vcross(q.v, v), 2.0f);
    return vadd(vadd(v, vscale(t, q.w)), vcross(q.v, t));
}
static quat qconj(quat q) { quat r = { q.w, vneg(q.v) }; return r; }

/* Transform::timeIndex, RMath.h:850-884 */
static size_t time_index(const float* times, size_t n, float time, float* out_t)
{
    size_t lower = 0, upper = n - 1;
    if (times[upper] <= time) lower = upper;
    else if (times[lower] >= time) upper = lower;
    while (upper - lower > 0)
    {
        size_t mid = (lower + upper) / 2;
        if (time < times[mid]) upper = mid;
        else if (mid > lower) lower = mid;
        else break;
    }
    if (lower == n - 1) *out_t = 0.0f;
    else if (times[lower] >= time) *out_t = 0.0f;
    else *out_t = (time - times[lower]) / (times[lower + 1] - times[lower]);
    return lower;
}

static vec3 xf_translation(const RtSceneDesc* d, uint32_t xf, float time)
{
    const RtXform* x = &d->xforms[xf];
    if (x->num_keys == 0) return v3(0.0f, 0.0f, 0.0f);
    float t;
    size_t i = x->first_key + time_index(d->key_time + x->first_key, x->num_keys, time, &t);
    const float* k = d->key_translation + 3 * i;
    if (t == 0.0f) return v3(k[0], k[1], k[2]);
    return vadd(vscale(v3(k[0], k[1], k[2]), 1.0f - t), vscale(v3(k[3], k[4], k[5]), t));
}
static vec3 xf_scaling(const RtSceneDesc* d, uint32_t xf, float time)
{
    const RtXform* x = &d->xforms[xf];
    if (x->num_keys == 0) return v3(1.0f, 1.0f, 1.0f);
    float t;
    size_t i = x->first_key + time_index(d->key_time + x->first_key, x->num_keys, time, &t);
    const float* k = d->key_scale + 3 * i;
    if (t == 0.0f) return v3(k[0], k[1], k[2]);
    return vadd(vscale(v3(k[0], k[1], k[2]), 1.0f - t), vscale(v3(k[3], k[4], k[5]), t));
}
static quat xf_rotation(const RtSceneDesc* d, uint32_t xf, float time)
{
    const RtXform* x = &d->xforms[xf];
    quat q = { 1.0f, { 0.0f, 0.0f, 0.0f } };
    if (x->num_keys == 0) return q;
    float t;
    size_t i = x->first_key + time_index(d->key_time + x->first_key, x->num_keys, time, &t);
    const float* k = d->key_rotation + 4 * i;
    if (t == 0.0f) { q.w = k[0]; q.v = v3(k[1], k[2], k[3]); return q; }
    /* lerp(q1, q2, t) = (q1*(1-t) + q2*t).normalized() */
    float om = 1.0f - t;
    q.w = om * k[0] + t * k[4];
    q.v = vadd(vscale(v3(k[1], k[2], k[3]), om), vscale(v3(k[5], k[6], k[7]), t));
    float len = sqrtf(q.w * q.w + vlen2(q.v));
    if (len > 0) { q.w /= len; q.v.x /= len; q.v.y /= len; q.v.z /= len; }
    return q;
}
/* each accessor re-evaluated per use, as the reference does (RMath.h:814-842) */
static vec3 to_local_point(const RtSceneDesc* d, uint32_t xf, float time, vec3 p)
{
    return vdiv(qrot(qconj(xf_rotation(d, xf, time)), vsub(p, xf_translation(d, xf, time))), xf_scaling(d, xf, time));
}
static vec3 to_local_vector(const RtSceneDesc* d, uint32_t xf, float time, vec3 v)
{
    return vdiv(qrot(qconj(xf_rotation(d, xf, time)), v), xf_scaling(d, xf, time));
}
static vec3 from_local_normal(const RtSceneDesc* d, uint32_t xf, float time, vec3 n)
{
    return qrot(xf_rotation(d, xf, time), n);
}

/* ---- Ray / Intersection, RRay.h ----------------------------------------- */
typedef struct { vec3 o, d; float tmax, time; } ray_t;
typedef struct
{
    ray_t ray;
    float t;
    int shape, face, tri;
    vec3 normal;
    float color_mod;
} isect_t;

static ray_t ray_to_local(const RtSceneDesc* d, uint32_t xf, ray_t r)
{
    ray_t l = r;
    g_work.xform_evals++;
    if (d->xforms[xf].num_keys >= 1) g_work.xform_keyed++;
    if (d->xforms[xf].num_keys >= 2) g_work.xform_pairs++;
    l.o = to_local_point(d, xf, r.time, r.o);
    l.d = to_local_vector(d, xf, r.time, r.d);
    return l;
}
static vec3 ray_at(ray_t r, float t) { return vadd(r.o, vscale(r.d, t)); }

/* ---- BBox::intersects, RAccel.h:47-59 ------------------------------------ */
static int box_hit(const RtBvhNode* n, vec3 o, vec3 inv, float* t0, float* t1)
{
    vec3 lo = v3(n->bbox_min[0], n->bbox_min[1], n->bbox_min[2]);
    vec3 hi = v3(n->bbox_max[0], n->bbox_max[1], n->bbox_max[2]);
    vec3 a = vmul(vsub(lo, o), inv);
    vec3 b = vmul(vsub(hi, o), inv);
    vec3 nr = v3(fmin_std(a.x, b.x), fmin_std(a.y, b.y), fmin_std(a.z, b.z));
    vec3 fr = v3(fmax_std(a.x, b.x), fmax_std(a.y, b.y), fmax_std(a.z, b.z));
    float bt_min = fmax_std(fmax_std(nr.x, nr.y), nr.z);
    float bt_max = fmin_std(fmin_std(fr.x, fr.y), fr.z);
    *t0 = fmax_std(bt_min, *t0);
    *t1 = fmin_std(bt_max, *t1);
    return *t0 <= *t1;
}

/* ---- Mesh faces, RMesh.h:226-379 ------------------------------------------ */
static vec3 mesh_vertex(const RtSceneDesc* d, const RtMesh* m, uint32_t i)
{
    const float* p = d->vertices + 3 * (size_t)(m->first_vertex + i);
    return v3(p[0], p[1], p[2]);
}
static vec3 mesh_normal(const RtSceneDesc* d, const RtMesh* m, uint32_t i)
{
    const float* p = d->normals + 3 * (size_t)(m->first_normal + i);
    return v3(p[0], p[1], p[2]);
}

/* intersectTri (closest != 0) / doesIntersectTri; tlimit is m_t or ray.m_tMax */
static int tri_test(const RtSceneDesc* d, const RtMesh* m, uint32_t face, uint32_t tri, ray_t r, float tlimit,
                    int closest, isect_t* is)
{
    uint32_t gf = m->first_face + face;
    const uint32_t* vi = d->vertex_index + d->face_start[gf];
    g_work.tri_tests++;
    vec3 p0 = mesh_vertex(d, m, vi[0]), p1 = mesh_vertex(d, m, vi[tri + 1]), p2 = mesh_vertex(d, m, vi[tri + 2]);
    vec3 e1 = vsub(p1, p0), e2 = vsub(p2, p0);
    vec3 g = vcross(e1, e2);
    float det = -vdot(r.d, g);
    if (det == 0.0f) return 0;
    vec3 r0 = vsub(p0, r.o);
    vec3 rvc = vcross(r.d, r0);
    vec3 r1 = vsub(p1, r.o);
    float inv_det = 1.0f / det;
    float gamma = -vdot(r1, rvc) * inv_det;
    if (gamma < 0.0f || gamma > 1.0f) return 0;
    vec3 r2 = vsub(p2, r.o);
    float beta = vdot(r2, rvc) * inv_det;
    if (beta < 0.0f || beta + gamma > 1.0f) return 0;
    float t = -vdot(r0, g) * inv_det;
    if (t < K_RAY_TMIN || t >= tlimit) return 0;
    if (!closest) return 1;
    float alpha = 1.0f - beta - gamma;
    vec3 sn;
    if (d->face_has_normals[gf])
    {
        const uint32_t* ni = d->normal_index + d->face_start[gf];
        vec3 n0 = mesh_normal(d, m, ni[0]), n1 = mesh_normal(d, m, ni[tri + 1]), n2 = mesh_normal(d, m, ni[tri + 2]);
        sn = vnormalized(vadd(vadd(vscale(n0, alpha), vscale(n1, beta)), vscale(n2, gamma)));
    }
    else
        sn = vnormalized(g);
    is->t = t;
    is->face = (int)face;
    is->tri = (int)tri;
    is->normal = sn;
    is->color_mod = 1.0f;
    return 1;
}

/* Bvh<Mesh>::intersect / doesIntersect, RAccel.h:389-563, with Mesh's face hooks */
static int mesh_bvh(const RtSceneDesc* d, const RtMesh* m, ray_t r, int closest, isect_t* is)
{
    vec3 inv = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int sign[3] = { inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f };
    struct { uint32_t node; float t0, t1; } steps[K_MAX_STEPS + 1];
    unsigned num = m->num_nodes > 0 ? 1 : 0;
    const RtBvhNode* nodes = d->mesh_nodes + m->first_node;
    steps[0].node = 0;
    steps[0].t0 = K_RAY_TMIN;
    steps[0].t1 = closest ? is->t : r.tmax;
    int hit = 0;
    while (num > 0 && num <= K_MAX_STEPS)
    {
        unsigned s = num - 1;
        const RtBvhNode* n = &nodes[steps[s].node];
        g_work.node_pops++;
        if (num > g_work.max_stack_mesh) g_work.max_stack_mesh = num;
        if (n->flags & 4u)
        {
            uint32_t face = n->first_child_or_prim;
            uint32_t gf = m->first_face + face;
            uint32_t nv = d->face_start[gf + 1] - d->face_start[gf];
            for (uint32_t k = 0; k + 2 < nv; ++k)
            {
                if (tri_test(d, m, face, k, r, closest ? is->t : r.tmax, closest, is))
                {
                    if (!closest) return 1;
                    hit = 1;
                }
            }
            num--;
            continue;
        }
        float t0 = steps[s].t0, t1 = steps[s].t1;
        if (closest)
        {
            if (t0 >= is->t) { num--; continue; }
            if (t1 > is->t) t1 = is->t;
        }
        if (!box_hit(n, r.o, inv, &t0, &t1)) { num--; continue; }
        uint32_t closest_node, furthest_node;
        if (!sign[n->flags & 3u]) { furthest_node = n->first_child_or_prim; closest_node = n->first_child_or_prim + 1; }
        else { closest_node = n->first_child_or_prim; furthest_node = n->first_child_or_prim + 1; }
        steps[s].node = furthest_node; steps[s].t0 = t0; steps[s].t1 = t1;
        num++; s++;
        steps[s].node = closest_node; steps[s].t0 = t0; steps[s].t1 = t1;
    }
    return hit;
}

/* ---- Shape::intersect / doesIntersect per type ----------------------------- */
/* ray is the set-local ray; returns 1 on (closer) hit.  closest == 0: any-hit rules */
static int shape_test(const RtSceneDesc* d, uint32_t sid, ray_t ray, int closest, isect_t* is)
{
    const RtShape* sh = &d->shapes[sid];
    float tlimit = closest ? is->t : ray.tmax;
    ray_t l = ray_to_local(d, sh->xform, ray);
    if (sh->type == RT_SHAPE_PLANE)
    {
        /* Plane::intersect / doesIntersect, RScene.h:288-363 */
        const RtPlane* p = &d->planes[sh->geom];
        g_work.shape_tests++;
        vec3 n = v3(p->normal[0], p->normal[1], p->normal[2]);
        vec3 pos = v3(p->position[0], p->position[1], p->position[2]);
        float ndd = vdot(n, l.d);
        if (ndd >= 0.0f) return 0;
        float t = (vdot(pos, n) - vdot(l.o, n)) / ndd;
        if (t >= tlimit || t < K_RAY_TMIN) return 0;
        if (!closest) return 1;
        is->t = t;
        is->shape = (int)sid; is->face = is->tri = -1;
        is->normal = from_local_normal(d, sh->xform, l.time, n);
        is->color_mod = 1.0f;
        if (p->bullseye && fmodf(vlen(vsub(ray_at(l, t), pos)) * 0.25f, 1.0f) > 0.5f)
            is->color_mod *= 0.2f;
        return 1;
    }
    if (sh->type == RT_SHAPE_SPHERE)
    {
        /* Sphere::intersect RScene.h:397-466 / doesIntersect :468-512 */
        const RtSphere* s = &d->spheres[sh->geom];
        g_work.shape_tests++;
        l.o = vsub(l.o, v3(s->position[0], s->position[1], s->position[2]));
        float a = vlen2(l.d);
        float b = 2.0f * vdot(l.d, l.o);
        float c = vlen2(l.o) - s->radius * s->radius;
        float disc = b * b - 4.0f * a * c;
        if (disc < 0.0f) return 0;
        disc = sqrtf(disc);
        float q = (b < 0.0f) ? (-0.5f * (b - disc)) : (-0.5f * (b + disc));
        float t0 = q / a;
        if (!closest)
        {
            if (t0 >= K_RAY_TMIN && t0 < ray.tmax) return 1;
            float t1 = c / q;
            if (q != 0.0f && t1 < ray.tmax && t1 >= K_RAY_TMIN) return 1;
            return 0;
        }
        float t1 = (q != 0.0f) ? (c / q) : is->t;
        if (t0 > t1) { float tmp = t1; t1 = t0; t0 = tmp; }
        if (t0 >= K_RAY_TMIN && t0 < is->t) is->t = t0;
        else if (t1 >= K_RAY_TMIN && t1 < is->t) is->t = t1;
        else return 0;
        is->shape = (int)sid; is->face = is->tri = -1;
        is->normal = vnormalized(from_local_normal(d, sh->xform, l.time, ray_at(l, is->t)));
        is->color_mod = 1.0f;
        return 1;
    }
    if (sh->type == RT_SHAPE_RECT)
    {
        /* RectangleLight::intersect RLight.h:58-116 / doesIntersect :118-163 */
        const RtRect* rc = &d->rects[sh->geom];
        g_work.shape_tests++;
        vec3 pos = v3(rc->position[0], rc->position[1], rc->position[2]);
        vec3 s1 = v3(rc->side1[0], rc->side1[1], rc->side1[2]);
        vec3 s2 = v3(rc->side2[0], rc->side2[1], rc->side2[2]);
        vec3 n = vnormalized(vcross(s1, s2));
        float ndd = vdot(n, l.d);
        if (ndd == 0.0f) return 0;
        float t = (vdot(pos, n) - vdot(l.o, n)) / ndd;
        if (t >= tlimit || t < K_RAY_TMIN) return 0;
        float len1 = vlen(s1), len2 = vlen(s2);
        vec3 s1n = vnormalized(s1), s2n = vnormalized(s2);
        vec3 rel = vsub(ray_at(l, t), pos);
        float u = vdot(rel, s1n), v = vdot(rel, s2n);
        if (u < 0.0f || u > len1 || v < 0.0f || v > len2) return 0;
        if (!closest) return 1;
        is->t = t;
        is->shape = (int)sid; is->face = is->tri = -1;
        is->color_mod = 1.0f;
        is->normal = from_local_normal(d, sh->xform, l.time, n);
        if (vdot(is->normal, ray.d) > 0.0f)
            is->normal = vscale(is->normal, -1.0f);
        return 1;
    }
    /* Mesh::intersect RMesh.h:62-74 / doesIntersect :76-81 */
    {
        const RtMesh* m = &d->meshes[sh->geom];
        if (!closest)
            return mesh_bvh(d, m, l, 0, is);
        isect_t local = *is;
        if (!mesh_bvh(d, m, l, 1, &local))
            return 0;
        *is = local;
        is->shape = (int)sid;
        is->normal = from_local_normal(d, sh->xform, ray.time, is->normal);
        return 1;
    }
}

/* Bvh<ShapeSet>::intersect / doesIntersect over the finite shapes */
static int top_bvh(const RtSceneDesc* d, ray_t r, int closest, isect_t* is)
{
    vec3 inv = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    int sign[3] = { inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f };
    struct { uint32_t node; float t0, t1; } steps[K_MAX_STEPS + 1];
    unsigned num = d->num_top_nodes > 0 ? 1 : 0;
    steps[0].node = 0;
    steps[0].t0 = K_RAY_TMIN;
    steps[0].t1 = closest ? is->t : r.tmax;
    int hit = 0;
    while (num > 0 && num <= K_MAX_STEPS)
    {
        unsigned s = num - 1;
        const RtBvhNode* n = &d->top_nodes[steps[s].node];
        g_work.node_pops++;
        if (num > g_work.max_stack_top) g_work.max_stack_top = num;
        if (n->flags & 4u)
        {
            if (shape_test(d, n->first_child_or_prim, r, closest, is))
            {
                if (!closest) return 1;
                hit = 1;
            }
            num--;
            continue;
        }
        float t0 = steps[s].t0, t1 = steps[s].t1;
        if (closest)
        {
            if (t0 >= is->t) { num--; continue; }
            if (t1 > is->t) t1 = is->t;
        }
        if (!box_hit(n, r.o, inv, &t0, &t1)) { num--; continue; }
        uint32_t closest_node, furthest_node;
        if (!sign[n->flags & 3u]) { furthest_node = n->first_child_or_prim; closest_node = n->first_child_or_prim + 1; }
        else { closest_node = n->first_child_or_prim; furthest_node = n->first_child_or_prim + 1; }
        steps[s].node = furthest_node; steps[s].t0 = t0; steps[s].t1 = t1;
        num++; s++;
        steps[s].node = closest_node; steps[s].t0 = t0; steps[s].t1 = t1;
    }
    return hit;
}

/* ShapeSet::intersect, RScene.h:120-156 */
static int set_intersect(const RtSceneDesc* d, ray_t world, isect_t* is)
{
    ray_t r = ray_to_local(d, d->set_xform, world);
    is->ray = r;
    is->t = world.tmax;
    is->shape = is->face = is->tri = -1;
    is->normal = v3(0.0f, 0.0f, 0.0f);
    is->color_mod = 1.0f;
    int any = 0;
    for (uint32_t i = 0; i < d->num_infinite; ++i)
        if (shape_test(d, d->num_finite + i, r, 1, is)) any = 1;
    if (d->num_finite > 2)
    {
        if (top_bvh(d, r, 1, is)) any = 1;
    }
    else
    {
        for (uint32_t i = 0; i < d->num_finite; ++i)
            if (shape_test(d, i, r, 1, is)) any = 1;
    }
    if (any)
        is->normal = from_local_normal(d, d->set_xform, world.time, is->normal);
    return any;
}

/* ShapeSet::doesIntersect, RScene.h:158-184 */
static int set_does_intersect(const RtSceneDesc* d, ray_t world)
{
    ray_t r = ray_to_local(d, d->set_xform, world);
    isect_t dummy;
    memset(&dummy, 0, sizeof(dummy));
    for (uint32_t i = 0; i < d->num_infinite; ++i)
        if (shape_test(d, d->num_finite + i, r, 0, &dummy)) return 1;
    if (d->num_finite > 2)
        return top_bvh(d, r, 0, &dummy);
    for (uint32_t i = 0; i < d->num_finite; ++i)
        if (shape_test(d, i, r, 0, &dummy)) return 1;
    return 0;
}

void port_trace_closest(const RtSceneDesc* d, const RtRay* rays, size_t n, RtHitEx* out)
{
    for (size_t i = 0; i < n; ++i)
    {
        ray_t r;
        r.o = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        r.d = v3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        r.tmax = rays[i].tmax;
        r.time = rays[i].time;
        isect_t is;
        int hit = set_intersect(d, r, &is);
        out[i].t = is.t;
        out[i].shape = hit ? is.shape : -1;
        out[i].face = hit ? is.face : -1;
        out[i].tri = hit ? is.tri : -1;
        out[i].normal[0] = is.normal.x; out[i].normal[1] = is.normal.y; out[i].normal[2] = is.normal.z;
        out[i].color_modifier = is.color_mod;
    }
}

void port_trace_any(const RtSceneDesc* d, const RtRay* rays, size_t n, uint8_t* out)
{
    for (size_t i = 0; i < n; ++i)
    {
        ray_t r;
        r.o = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        r.d = v3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        r.tmax = rays[i].tmax;
        r.time = rays[i].time;
        out[i] = (uint8_t)set_does_intersect(d, r);
    }
}

/* ---- sample stream ---------------------------------------------------------- */
/* Rng::nextUInt32, RSampling.h:52-57 */
static uint32_t rng_next(uint32_t* z, uint32_t* w)
{
    *z = 36969u * (*z & 65535u) + (*z >> 16);
    *w = 18000u * (*w & 65535u) + (*w >> 16);
    return (*z << 16) + *w;
}

/* The permutations pixel (x, y) renders with, by walking the chunk's Rng LITERALLY
 * through every earlier pixel as RenderThread::run does (RaytraceMain.cpp:69-108,
 * 112-169).  Layout of out: see rt_sample_permutations in rayito_b200.h.
 * Returns 0, or -1 when the pixel lies in no chunk (images narrower than 4). */
int port_pixel_permutations(uint32_t width, uint32_t height, uint32_t depth, uint32_t px, uint32_t py, uint32_t* out)
{
    size_t cw = width >= 4 ? width / 4 : 1, ch = height >= 4 ? height / 4 : 1;      /* :508-509 */
    size_t nx = width > 4 ? width / cw : 1, ny = height > 4 ? height / ch : 1;      /* :513-514 */
    if (nx * cw < width) nx++;
    if (ny * ch < height) ny++;
    size_t cx = px / cw, cy = py / ch;
    if (cx >= nx || cy >= ny) return -1;
    size_t xs = cx * cw, ys = cy * ch;
    size_t xe = (cx + 1) * cw < width ? (cx + 1) * cw : width;
    size_t ye = (cy + 1) * ch < height ? (cy + 1) * ch : height;
    uint32_t z = (uint32_t)(((xs << 16) | xe) ^ xs), w = (uint32_t)(((ys << 16) | ye) ^ ys);   /* :69-70 */
    uint32_t perm[5 * 16 + 3];
    uint32_t per = 5 * depth + 3;
    /* construction: per bounce {bounce, lightSel, lightElem, light, brdf}; then time, lens, subpixel */
    for (uint32_t i = 0; i < 5 * depth; ++i) perm[i] = rng_next(&z, &w);
    perm[5 * depth + 0] = rng_next(&z, &w);   /* time */
    perm[5 * depth + 1] = rng_next(&z, &w);   /* lens */
    perm[5 * depth + 2] = rng_next(&z, &w);   /* subpixel */
    for (size_t y = ys; y < ye; ++y)
        for (size_t x = xs; x < xe; ++x)
        {
            if (x == px && y == py)
            {
                memcpy(out, perm, per * sizeof(uint32_t));
                return 0;
            }
            /* refill after the pixel: bounces, then lens, time, subpixel (:159-169) */
            for (uint32_t i = 0; i < 5 * depth; ++i) perm[i] = rng_next(&z, &w);
            perm[5 * depth + 1] = rng_next(&z, &w);   /* lens */
            perm[5 * depth + 0] = rng_next(&z, &w);   /* time */
            perm[5 * depth + 2] = rng_next(&z, &w);   /* subpixel */
        }
    return -1;
}

/* CorrelatedMultiJitterSampler::permute / randFloat01, RSampling.h:328-374 */
static uint32_t cmj_permute(uint32_t i, uint32_t num, uint32_t p)
{
    uint32_t w = num - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do
    {
        i ^= p;             i *= 0xe170893du;
        i ^= p >> 16;       i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3fu;
        i ^= p >> 23;       i ^= (i & w) >> 1;
        i *= 1u | p >> 27;  i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;  i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= num);
    return (i + p) % num;
}
static float cmj_rand(uint32_t i, uint32_t p)
{
    i ^= p;
    i ^= i >> 17; i ^= i >> 10; i *= 0xb36534e5u;
    i ^= i >> 12; i ^= i >> 21; i *= 0x93fc4795u;
    i ^= 0xdf6e307fu;
    i ^= i >> 17; i *= 1u | p >> 18;
    return i * 2.328306e-10f;
}
float port_cmj_1d(uint32_t index, uint32_t samples, uint32_t perm)     /* RSampling.h:272-279 */
{
    uint32_t s = cmj_permute(index, samples, perm * 0x8ff3cd11u);
    float sx = cmj_rand(s, perm * 0xa399d265u);
    return (s + sx) / (float)samples;
}
void port_cmj_2d(uint32_t index, uint32_t xs, uint32_t ys, uint32_t perm, float* u, float* v)   /* :288-306 */
{
    uint32_t s = cmj_permute(index, xs * ys, perm * 0xc2d3c8fbu);
    int ix = (int)cmj_permute(s % xs, xs, perm * 0xa511e9b3u);
    int iy = (int)cmj_permute(s / xs, ys, perm * 0x63d83595u);
    float sx = cmj_rand(s, perm * 0xa399d265u);
    float sy = cmj_rand(s, perm * 0x711ad6a5u);
    *u = (ix + (iy + sx) / (float)ys) / (float)xs;
    *v = (s + sy) / (float)(xs * ys);
}

/* RenderThread::run sample set-up + PerspectiveCamera::makeRay (pinhole),
 * RaytraceMain.cpp:120-142, 224-236 */
void port_camera_ray(const RtCamera* cam, uint32_t width, uint32_t height, uint32_t ps, uint32_t depth,
                     uint32_t x, uint32_t y, uint32_t psi, RtRay* out)
{
    uint32_t perm[5 * 16 + 3];
    memset(out, 0, sizeof(*out));
    if (port_pixel_permutations(width, height, depth, x, y, perm) != 0)
        return;
    float pu, pv, lu, lv;
    port_cmj_2d(psi, ps, ps, perm[5 * depth + 2], &pu, &pv);
    float xu = (x + pu) / (float)width;
    float yu = 1.0f - (y + pv) / (float)height;
    port_cmj_2d(psi, ps, ps, perm[5 * depth + 1], &lu, &lv);
    float time_u = port_cmj_1d(psi, ps * ps, perm[5 * depth + 0]);
    float aspect = (float)width / (float)height;
    float xs = (xu - 0.5f) * aspect + 0.5f, ysc = yu;
    vec3 fwd = v3(cam->forward[0], cam->forward[1], cam->forward[2]);
    vec3 right = v3(cam->right[0], cam->right[1], cam->right[2]);
    vec3 up = v3(cam->up[0], cam->up[1], cam->up[2]);
    vec3 dir = vadd(vadd(fwd, vscale(right, (xs - 0.5f) * cam->tan_fov)), vscale(up, (ysc - 0.5f) * cam->tan_fov));
    dir = vnormalized(dir);
    out->origin[0] = cam->origin[0]; out->origin[1] = cam->origin[1]; out->origin[2] = cam->origin[2];
    out->direction[0] = dir.x; out->direction[1] = dir.y; out->direction[2] = dir.z;
    out->tmax = K_RAY_TMAX;
    out->time = cam->shutter_open + (cam->shutter_close - cam->shutter_open) * time_u;
}

// BRDFs, light sampling and light pdfs of the shading stage.
//
// Reference: Rayito_Stage7_QT/RMaterial.h (Lambert :91-204, Glossy :208-372,
// PerfectReflection :376-431), RLight.h (RectangleLight::sampleSurface :186-218,
// intersectPdf :220-239; ShapeLight :293-328), RScene.h (Sphere::sampleSurface
// :528-575, pdfSA :577-596, surfaceAreaPdf :598-601), RMesh.h (sampleSurface
// :133-185, pdfSA :187-196).
//
// Every expression that contains M_PI is evaluated in double and rounded once, as
// the CPU does (M_PI is a double literal).
#ifndef RAYITO_B200_RT_SHADE_CUH
#define RAYITO_B200_RT_SHADE_CUH

#include "rt_sampling.cuh"
#include "rt_trace.cuh"

// Inlined by default.  k_shade / k_light_sample / k_resolve come to 58-76 KB of SASS and show
// 30 % no-instruction stalls, but making these routines real calls (-DRT_SHADE_CALL=__noinline__)
// measured 5 % SLOWER on the frame (profiles/README.md, round 1): the calls' spills cost more
// than the fetch stalls they save.
#ifndef RT_SHADE_CALL
#define RT_SHADE_CALL __forceinline__
#endif

struct Color3
{
    float r, g, b;
};

__device__ __forceinline__ Color3 mkc(float r, float g, float b) { Color3 c; c.r = r; c.g = g; c.b = b; return c; }
__device__ __forceinline__ Color3 operator+(Color3 a, Color3 b) { return mkc(a.r + b.r, a.g + b.g, a.b + b.b); }
__device__ __forceinline__ Color3 operator*(Color3 a, Color3 b) { return mkc(a.r * b.r, a.g * b.g, a.b * b.b); }
__device__ __forceinline__ Color3 operator*(Color3 a, float f) { return mkc(f * a.r, f * a.g, f * a.b); }
__device__ __forceinline__ Color3 operator/(Color3 a, float f) { return mkc(a.r / f, a.g / f, a.b / f); }

// ---------------------------------------------------------------------------
// BRDFs.  "incoming" points TOWARDS the surface, "outgoing" away from it.
// ---------------------------------------------------------------------------

__device__ __forceinline__ bool same_hemisphere_reject(float n_dot_i, float n_dot_o)
{
    return (n_dot_i > 0.0f && n_dot_o > 0.0f) || (n_dot_i < 0.0f && n_dot_o < 0.0f);
}

// Brdf::evaluateSA; returns the reflectance, pdf through out_pdf
__device__ RT_SHADE_CALL float brdf_evaluate(uint32_t brdf, float exponent, V3 incoming, V3 outgoing, V3 normal, float& out_pdf)
{
    if (brdf == RT_BRDF_LAMBERT)
    {
        float n_dot_i = dot3(incoming, normal);
        float n_dot_o = dot3(outgoing, normal);
        if (same_hemisphere_reject(n_dot_i, n_dot_o))
        {
            out_pdf = 0.0f;
            return 0.0f;
        }
        out_pdf = (float)((double)fabsf(n_dot_i) / RT_PI_D);
        return (float)(1.0 / RT_PI_D);
    }
    if (brdf == RT_BRDF_GLOSSY)
    {
        float n_dot_i = dot3(incoming, normal);
        float n_dot_o = dot3(outgoing, normal);
        if (same_hemisphere_reject(n_dot_i, n_dot_o))
        {
            out_pdf = 0.0f;
            return 0.0f;
        }
        V3 half;
        if (dot3(outgoing, incoming) > 0.999f)
            half = normal;
        else
            half = normalized3(outgoing - incoming);
        // d = (e + 1) * pow(|n.h|, e) / (2 pi): float product, double division
        float d = (float)((double)((exponent + 1.0f) * ref_powf(fabsf(dot3(normal, half)), exponent)) / (2.0 * RT_PI_D));
        float result = 1.0f * d / (4.0f * fabsf(n_dot_o + -n_dot_i - n_dot_o * -n_dot_i));
        out_pdf = d / (4.0f * fabsf(dot3(outgoing, half)));
        return result;
    }
    // PerfectReflection (and Emitter, which never gets here)
    out_pdf = 0.0f;
    return 0.0f;
}

// Brdf::sampleSA
__device__ RT_SHADE_CALL float brdf_sample(uint32_t brdf, float exponent, V3& out_incoming, V3 outgoing, V3 normal,
                                             float u1, float u2, float& out_pdf)
{
    if (brdf == RT_BRDF_LAMBERT)
    {
        V3 local = -cosine_hemisphere(u1, u2);
        V3 x, y, z;
        make_frame(normal, x, y, z);
        out_incoming = frame_to_world(local, x, y, z);
        if (dot3(outgoing, normal) < 0.0f)
            out_incoming = out_incoming * -1.0f;
        out_pdf = (float)((double)fabsf(dot3(-out_incoming, normal)) / RT_PI_D);
        return (float)(1.0 / RT_PI_D);
    }
    if (brdf == RT_BRDF_GLOSSY)
    {
        float phi = (float)((2.0 * RT_PI_D) * (double)u1);      // 2.0f * M_PI * u1
        float cos_theta = ref_powf(1.0f - u2, 1.0f / (exponent + 1.0f));
        float sin2_theta = std_max(0.0f, 1.0f - cos_theta * cos_theta);
        float sin_theta = sqrtf(sin2_theta);
        float sn, cs;
        ref_sincosf(phi, sn, cs);
        V3 local_half = mk(sin_theta * cs, sin_theta * sn, cos_theta);
        V3 x, y, z;
        make_frame(normal, x, y, z);
        V3 half = frame_to_world(local_half, x, y, z);
        if (dot3(outgoing, normal) < 0.0f)
            half = half * -1.0f;
        out_incoming = outgoing - half * (2.0f * dot3(outgoing, half));
        return brdf_evaluate(RT_BRDF_GLOSSY, exponent, out_incoming, outgoing, normal, out_pdf);
    }
    // PerfectReflection (RMaterial.h:399-407)
    float n_dot_o = dot3(normal, outgoing);
    if (n_dot_o < 0.0f)
        out_incoming = outgoing + 2.0f * normal * n_dot_o;
    else
        out_incoming = outgoing - 2.0f * normal * n_dot_o;
    out_pdf = fabsf(dot3(-out_incoming, normal));
    return 1.0f;
}

// ---------------------------------------------------------------------------
// Lights
// ---------------------------------------------------------------------------

// Sphere::surfaceAreaPdf (RScene.h:598-601): 3 / (4 pi r^2) -- sic, in double
__device__ __forceinline__ float sphere_area_pdf(float radius)
{
    return (float)(3.0 / (((4.0 * RT_PI_D) * (double)radius) * (double)radius));
}

// Light::sampleSurface for the light behind shape `sh`.  Positions are in the
// space of the ShapeSet's members ("non-local"): the reference does not apply the
// set's own transform here either.
__device__ RT_SHADE_CALL void light_sample(const DScene& sc, const DShape& sh, V3 ref_pos, float ref_time,
                                             float u1, float u2, float u3,
                                             V3& out_pos, V3& out_normal, float& out_pdf, const float4* row = nullptr)
{
    out_pdf = 0.0f;
    out_pos = mk(0.0f, 0.0f, 0.0f);
    out_normal = mk(0.0f, 0.0f, 0.0f);
    TRS trs = shape_xform(sc, sh, ref_time, row);
    if (sh.type == RT_SHAPE_RECT)
    {
        // RLight.h:186-218
        DRect rc = sc.rects[sh.geom];
        V3 side1 = mk(rc.r1x, rc.r1y, rc.r1z), side2 = mk(rc.r2x, rc.r2y, rc.r2z);
        V3 p = mk(rc.px, rc.py, rc.pz) + side1 * u1 + side2 * u2;
        p = from_local_point(trs, p);
        V3 outgoing = ref_pos - p;
        float dist;
        outgoing = normalized3(outgoing, &dist);
        V3 n = cross3(side1, side2);
        n = from_local_vector(trs, n);
        float area;
        n = normalized3(n, &area);
        if (dot3(n, outgoing) < 0.0f)
            n = n * -1.0f;
        float pdf = dist * dist / (area * fabsf(dot3(n, outgoing)));
        out_pos = p;
        out_normal = n;
        out_pdf = pdf > 1.0e10f ? 0.0f : pdf;
        return;
    }
    if (sh.type == RT_SHAPE_SPHERE)
    {
        // ShapeLight -> Sphere::sampleSurface (RScene.h:528-575).  The back-side
        // "discard" of ShapeLight::sampleSurface (RLight.h:310-313) only changes a
        // return value the path tracer ignores, so it has no effect here either.
        DSphere s = sc.spheres[sh.geom];
        V3 centre = mk(s.px, s.py, s.pz);
        V3 local_ref = to_local_point(trs, ref_pos);
        V3 to_centre = centre - local_ref;
        float dist2 = length2(to_centre);
        if (dist2 < s.radius * s.radius * 1.00001f)
        {
            V3 n = uniform_sphere(u1, u2);
            V3 p = centre + n * s.radius;
            n = from_local_normal(trs, n);
            p = from_local_point(trs, p);
            V3 to_surf = ref_pos - p;
            out_pdf = length2(to_surf) * sphere_area_pdf(s.radius) / fabsf(dot3(normalized3(to_surf), n));
            out_pos = p;
            out_normal = n;
            return;
        }
        float sin_theta_max2 = s.radius * s.radius / dist2;
        float cos_theta_max = sqrtf(std_max(0.0f, 1.0f - sin_theta_max2));
        V3 x, y, z;
        make_frame(to_centre, x, y, z);
        V3 local_cone = uniform_cone(u1, u2, cos_theta_max);
        V3 cone = normalized3(frame_to_world(local_cone, x, y, z));
        // The probe ray is built with time 0 (Ray's default), pushed out of local
        // space and straight back in by Sphere::intersect -- at time 0, not at the
        // sample's time (RScene.h:562-565, RRay.h:57).  Reproduced literally.
        TRS trs0 = shape_xform(sc, sh, 0.0f);
        V3 wo = from_local_point(trs0, local_ref);
        V3 wd = from_local_vector(trs0, cone);
        V3 lo = to_local_point(trs0, wo) - centre;
        V3 ld = to_local_vector(trs0, wd);
        float t;
        if (!sphere_closest(lo, ld, s.radius, RT_RAY_TMAX, t))
            t = dot3(to_centre, cone);
        V3 p = local_ref + t * cone;
        V3 n = normalized3(p - centre);
        out_normal = from_local_normal(trs, n);
        out_pos = from_local_point(trs, p);
        out_pdf = uniform_cone_pdf(cos_theta_max);
        return;
    }
    if (sh.type == RT_SHAPE_MESH)
    {
        // ShapeLight -> Mesh::sampleSurface (RMesh.h:133-185)
        DMesh m = sc.meshes[sh.geom];
        if (m.num_faces == 0)
            return;
        const float* cdf = sc.face_area_cdf + m.first_cdf;
        float pick = u3 * m.total_area;
        // std::upper_bound over num_faces + 1 entries: first element > pick
        uint32_t lo = 0, count = m.num_faces + 1;
        while (count > 0)
        {
            uint32_t step = count / 2;
            uint32_t mid = lo + step;
            if (!(pick < cdf[mid])) { lo = mid + 1; count -= step + 1; }
            else count = step;
        }
        uint32_t face;
        if (lo == m.num_faces + 1) face = m.num_faces;      // sic: size() - 1 of the CDF (RMesh.h:151-152)
        else if (lo == 0) face = 0;
        else face = lo - 1;
        if (face >= m.num_faces)
            return;                                         // the reference would read past the face list here
        float face_area = cdf[face + 1] - cdf[face];
        float selector = (pick - cdf[face]) / face_area;
        // Triangles of this face: the leaf that holds it is not known here, but the
        // records of a mesh are laid out face by face, so locate them by scanning the
        // face's fan via the record words (v0.w = face).  first_tri + sum of earlier
        // fans is precomputed per face in tri_face_start (see rt_scene.cuh).
        uint32_t rec = sc.face_first_tri[m.first_face + face];
        uint32_t rec_end = sc.face_first_tri[m.first_face + face + 1];
        float so_far = 0.0f;
        for (; rec < rec_end; ++rec)
        {
            V3 p0, p1, p2;
            uint32_t w0, w1, w2;
            load_tri(sc, rec, p0, p1, p2, w0, w1, w2);
            so_far += length3(cross3(p1 - p0, p2 - p0)) * 0.5f;
            if (selector * face_area < so_far)
            {
                float alpha, beta;
                uniform_barycentric(u1, u2, alpha, beta);
                float gamma = 1.0f - alpha - beta;
                V3 p = p0 * alpha + p1 * beta + p2 * gamma;
                p = from_local_point(trs, p);
                V3 n = cross3(p1 - p0, p2 - p0);
                n = normalized3(from_local_normal(trs, n));
                V3 to_surf = ref_pos - p;
                out_pdf = length2(to_surf) * (1.0f / m.total_area) / fabsf(dot3(normalized3(to_surf), n));
                out_pos = p;
                out_normal = n;
                return;
            }
        }
    }
}

// Light::intersectPdf for a BRDF-sampled ray that hit the light (RLight.h:220-239,
// 317-328).  ray_o/ray_d/time describe the probe ray, t/normal its hit.
__device__ RT_SHADE_CALL float light_intersect_pdf(const DScene& sc, const DShape& sh, V3 ray_o, V3 ray_d, float time,
                                                     float t, V3 hit_normal, const float4* row = nullptr)
{
    TRS trs = shape_xform(sc, sh, time, row);
    if (sh.type == RT_SHAPE_RECT)
    {
        DRect rc = sc.rects[sh.geom];
        V3 side1 = from_local_vector(trs, mk(rc.r1x, rc.r1y, rc.r1z));
        V3 side2 = from_local_vector(trs, mk(rc.r2x, rc.r2y, rc.r2z));
        float pdf = t * t / (fabsf(dot3(hit_normal, -ray_d)) * length3(cross3(side1, side2)));
        return pdf > 1.0e10f ? 0.0f : pdf;
    }
    V3 surf_pos = ray_o + t * ray_d;        // isect.position()
    if (sh.type == RT_SHAPE_SPHERE)
    {
        // Sphere::pdfSA (RScene.h:577-596)
        DSphere s = sc.spheres[sh.geom];
        V3 local_ref = to_local_point(trs, ray_o);
        V3 to_centre = mk(s.px, s.py, s.pz) - local_ref;
        float dist2 = length2(to_centre);
        if (dist2 < s.radius * s.radius * 1.00001f)
        {
            V3 to_surf = ray_o - surf_pos;
            return length2(to_surf) * sphere_area_pdf(s.radius) / fabsf(dot3(normalized3(to_surf), hit_normal));
        }
        float sin_theta_max2 = s.radius * s.radius / dist2;
        float cos_theta_max = sqrtf(std_max(0.0f, 1.0f - sin_theta_max2));
        return uniform_cone_pdf(cos_theta_max);
    }
    if (sh.type == RT_SHAPE_MESH)
    {
        // Mesh::pdfSA (RMesh.h:187-196)
        DMesh m = sc.meshes[sh.geom];
        V3 to_surf = ray_o - surf_pos;
        return length2(to_surf) * (1.0f / m.total_area) / fabsf(dot3(normalized3(to_surf), hit_normal));
    }
    return 0.0f;
}

#endif // RAYITO_B200_RT_SHADE_CUH

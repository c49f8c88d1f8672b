// Host-side value types of the Rayito API surface: Color, Vector/Point,
// Quaternion and the keyed Transform.
//
// These mirror the public names of the reference's RMath.h so that scene-building
// code written for the tutorial compiles unchanged, but they exist here only to
// *describe* a scene: every hot-path evaluation happens on the GPU.  Where host
// arithmetic feeds data the GPU consumes (transform keys, bounding boxes for the
// BVH build) the float operations are kept in the same association as the
// reference so the flattened scene is bit-identical to what the reference holds
// after prepare().  Each such spot cites reference file:line (Rayito_Stage7_QT).
#ifndef RAYITO_B200_MATH_HPP
#define RAYITO_B200_MATH_HPP

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace Rayito
{

struct Color
{
    float m_r, m_g, m_b;

    Color() : m_r(0.0f), m_g(0.0f), m_b(0.0f) { }
    Color(float r, float g, float b) : m_r(r), m_g(g), m_b(b) { }
    explicit Color(float f) : m_r(f), m_g(f), m_b(f) { }

    // RMath.h:46-51: max(lo, min(hi, c)) with std::min/max argument order
    void clamp(float lo = 0.0f, float hi = 1.0f)
    {
        m_r = std::max(lo, std::min(hi, m_r));
        m_g = std::max(lo, std::min(hi, m_g));
        m_b = std::max(lo, std::min(hi, m_b));
    }

    Color& operator+=(const Color& c) { m_r += c.m_r; m_g += c.m_g; m_b += c.m_b; return *this; }
    Color& operator-=(const Color& c) { m_r -= c.m_r; m_g -= c.m_g; m_b -= c.m_b; return *this; }
    Color& operator*=(const Color& c) { m_r *= c.m_r; m_g *= c.m_g; m_b *= c.m_b; return *this; }
    Color& operator/=(const Color& c) { m_r /= c.m_r; m_g /= c.m_g; m_b /= c.m_b; return *this; }
    Color& operator*=(float f) { m_r *= f; m_g *= f; m_b *= f; return *this; }
    Color& operator/=(float f) { m_r /= f; m_g /= f; m_b /= f; return *this; }
};

inline Color operator+(const Color& a, const Color& b) { return Color(a.m_r + b.m_r, a.m_g + b.m_g, a.m_b + b.m_b); }
inline Color operator-(const Color& a, const Color& b) { return Color(a.m_r - b.m_r, a.m_g - b.m_g, a.m_b - b.m_b); }
inline Color operator*(const Color& a, const Color& b) { return Color(a.m_r * b.m_r, a.m_g * b.m_g, a.m_b * b.m_b); }
inline Color operator/(const Color& a, const Color& b) { return Color(a.m_r / b.m_r, a.m_g / b.m_g, a.m_b / b.m_b); }
inline Color operator*(const Color& c, float f) { return Color(f * c.m_r, f * c.m_g, f * c.m_b); }
inline Color operator*(float f, const Color& c) { return Color(f * c.m_r, f * c.m_g, f * c.m_b); }
inline Color operator/(const Color& c, float f) { return Color(c.m_r / f, c.m_g / f, c.m_b / f); }


struct Vector
{
    float m_x, m_y, m_z;

    Vector() : m_x(0.0f), m_y(0.0f), m_z(0.0f) { }
    Vector(float x, float y, float z) : m_x(x), m_y(y), m_z(z) { }
    explicit Vector(float f) : m_x(f), m_y(f), m_z(f) { }

    // (x*x + y*y) + z*z, RMath.h:190
    float length2() const { return m_x * m_x + m_y * m_y + m_z * m_z; }
    float length() const { return std::sqrt(length2()); }

    // Divides (never multiplies by a reciprocal), only when len > 0: RMath.h:194
    float normalize()
    {
        float len = length();
        if (len > 0) { m_x /= len; m_y /= len; m_z /= len; }
        return len;
    }
    Vector normalized() const { Vector r(*this); r.normalize(); return r; }

    float maxComponent() const { return std::max(std::max(m_x, m_y), m_z); }
    float minComponent() const { return std::min(std::min(m_x, m_y), m_z); }

    Vector& operator+=(const Vector& v) { m_x += v.m_x; m_y += v.m_y; m_z += v.m_z; return *this; }
    Vector& operator-=(const Vector& v) { m_x -= v.m_x; m_y -= v.m_y; m_z -= v.m_z; return *this; }
    Vector& operator*=(const Vector& v) { m_x *= v.m_x; m_y *= v.m_y; m_z *= v.m_z; return *this; }
    Vector& operator/=(const Vector& v) { m_x /= v.m_x; m_y /= v.m_y; m_z /= v.m_z; return *this; }
    Vector& operator*=(float f) { m_x *= f; m_y *= f; m_z *= f; return *this; }
    Vector& operator/=(float f) { m_x /= f; m_y /= f; m_z /= f; return *this; }
    Vector operator-() const { return Vector(-m_x, -m_y, -m_z); }
};

typedef Vector Point;

inline Vector operator+(const Vector& a, const Vector& b) { return Vector(a.m_x + b.m_x, a.m_y + b.m_y, a.m_z + b.m_z); }
inline Vector operator-(const Vector& a, const Vector& b) { return Vector(a.m_x - b.m_x, a.m_y - b.m_y, a.m_z - b.m_z); }
inline Vector operator*(const Vector& a, const Vector& b) { return Vector(a.m_x * b.m_x, a.m_y * b.m_y, a.m_z * b.m_z); }
inline Vector operator/(const Vector& a, const Vector& b) { return Vector(a.m_x / b.m_x, a.m_y / b.m_y, a.m_z / b.m_z); }
inline Vector operator*(const Vector& v, float f) { return Vector(f * v.m_x, f * v.m_y, f * v.m_z); }
inline Vector operator*(float f, const Vector& v) { return Vector(f * v.m_x, f * v.m_y, f * v.m_z); }
inline Vector operator/(float f, const Vector& v) { return Vector(f / v.m_x, f / v.m_y, f / v.m_z); }
inline Vector operator/(const Vector& v, float f) { return Vector(v.m_x / f, v.m_y / f, v.m_z / f); }

inline float dot(const Vector& a, const Vector& b) { return a.m_x * b.m_x + a.m_y * b.m_y + a.m_z * b.m_z; }

inline Vector cross(const Vector& a, const Vector& b)
{
    return Vector(a.m_y * b.m_z - a.m_z * b.m_y,
                  a.m_z * b.m_x - a.m_x * b.m_z,
                  a.m_x * b.m_y - a.m_y * b.m_x);
}

// Component-wise std::max / std::min (first argument wins on NaN): RMath.h:348-360
inline Vector max(const Vector& a, const Vector& b)
{
    return Vector(std::max(a.m_x, b.m_x), std::max(a.m_y, b.m_y), std::max(a.m_z, b.m_z));
}
inline Vector min(const Vector& a, const Vector& b)
{
    return Vector(std::min(a.m_x, b.m_x), std::min(a.m_y, b.m_y), std::min(a.m_z, b.m_z));
}


struct Quaternion
{
    float m_w;
    Vector m_v;

    Quaternion() : m_w(1.0f), m_v(0.0f) { }
    Quaternion(float w, float x, float y, float z) : m_w(w), m_v(x, y, z) { }
    Quaternion(float w, const Vector& v) : m_w(w), m_v(v) { }
    // Axis + angle, float cos/sin of the half angle: RMath.h:395-396
    Quaternion(const Vector& axis, float angle)
        : m_w(std::cos(angle * 0.5f)), m_v(axis * std::sin(angle * 0.5f)) { }

    float length2() const { return m_w * m_w + m_v.length2(); }
    float length() const { return std::sqrt(length2()); }
    float normalize()
    {
        float len = length();
        if (len > 0) { m_w /= len; m_v /= len; }
        return len;
    }
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }

    Quaternion operator-() const { return Quaternion(-m_w, -m_v); }
    Quaternion operator~() const { return Quaternion(m_w, -m_v); }

    Quaternion& operator+=(const Quaternion& q) { m_w += q.m_w; m_v += q.m_v; return *this; }
    Quaternion& operator-=(const Quaternion& q) { m_w -= q.m_w; m_v -= q.m_v; return *this; }
    Quaternion& operator*=(float f) { m_w *= f; m_v *= f; return *this; }
    Quaternion& operator/=(float f) { m_w /= f; m_v /= f; return *this; }

    // In-place composition.  PARITY QUIRK (RMath.h:461-469): each component is
    // overwritten before the following lines read it, so this is NOT the Hamilton
    // product; Transform::rotate() keys are built through it and must match.
    Quaternion& operator*=(const Quaternion& q)
    {
        m_w     = m_w * q.m_w     - m_v.m_x * q.m_v.m_x - m_v.m_y * q.m_v.m_y - m_v.m_z * q.m_v.m_z;
        m_v.m_x = m_w * q.m_v.m_x + m_v.m_x * q.m_w     + m_v.m_y * q.m_v.m_z - m_v.m_z * q.m_v.m_y;
        m_v.m_y = m_w * q.m_v.m_y - m_v.m_x * q.m_v.m_z + m_v.m_y * q.m_w     + m_v.m_z * q.m_v.m_x;
        m_v.m_z = m_w * q.m_v.m_z + m_v.m_x * q.m_v.m_y - m_v.m_y * q.m_v.m_x + m_v.m_z * q.m_w;
        return *this;
    }
};

inline Quaternion operator+(const Quaternion& a, const Quaternion& b) { return Quaternion(a.m_w + b.m_w, a.m_v + b.m_v); }
inline Quaternion operator-(const Quaternion& a, const Quaternion& b) { return Quaternion(a.m_w - b.m_w, a.m_v - b.m_v); }
inline Quaternion operator*(const Quaternion& q, float f) { return Quaternion(f * q.m_w, f * q.m_v); }
inline Quaternion operator*(float f, const Quaternion& q) { return Quaternion(f * q.m_w, f * q.m_v); }

// Out-of-place Hamilton product (RMath.h:515-522)
inline Quaternion operator*(const Quaternion& a, const Quaternion& b)
{
    return Quaternion(a.m_w * b.m_w     - a.m_v.m_x * b.m_v.m_x - a.m_v.m_y * b.m_v.m_y - a.m_v.m_z * b.m_v.m_z,
                      a.m_w * b.m_v.m_x + a.m_v.m_x * b.m_w     + a.m_v.m_y * b.m_v.m_z - a.m_v.m_z * b.m_v.m_y,
                      a.m_w * b.m_v.m_y - a.m_v.m_x * b.m_v.m_z + a.m_v.m_y * b.m_w     + a.m_v.m_z * b.m_v.m_x,
                      a.m_w * b.m_v.m_z + a.m_v.m_x * b.m_v.m_y - a.m_v.m_y * b.m_v.m_x + a.m_v.m_z * b.m_w);
}

// Rotate a vector: t = 2 cross(qv, v); v + w t + cross(qv, t)  (RMath.h:536-549)
inline Vector operator*(const Quaternion& q, const Vector& v)
{
    Vector t = 2.0f * cross(q.m_v, v);
    return v + t * q.m_w + cross(q.m_v, t);
}

inline float dot(const Quaternion& a, const Quaternion& b) { return a.m_w * b.m_w + dot(a.m_v, b.m_v); }

// Normalised linear interpolation (RMath.h:576-580)
inline Quaternion lerp(const Quaternion& a, const Quaternion& b, float t)
{
    return (a * (1.0f - t) + b * t).normalized();
}


// Scale, then rotate, then translate; each animated by keys at strictly
// increasing times (RMath.h:619-941).
class Transform
{
public:
    Transform() { }

    size_t numKeys() const { return m_time.empty() ? 1 : m_time.size(); }
    size_t numSegments() const { return m_time.size() < 1 ? 0 : m_time.size() - 1; }
    float keyTime(size_t k) const { return k < m_time.size() ? m_time[k] : 0.0f; }
    // Number of stored keys; 0 for the keyless identity (unlike numKeys())
    size_t storedKeys() const { return m_time.size(); }

    void clear() { m_time.clear(); m_scale.clear(); m_rotate.clear(); m_translate.clear(); }

    Vector translationKey(size_t k) const
    {
        if (m_translate.empty()) return Vector(0.0f);
        return m_translate[std::min(k, m_translate.size() - 1)];
    }
    Vector scalingKey(size_t k) const
    {
        if (m_scale.empty()) return Vector(1.0f);
        return m_scale[std::min(k, m_scale.size() - 1)];
    }
    Quaternion rotationKey(size_t k) const
    {
        if (m_rotate.empty()) return Quaternion(1.0f, 0.0f, 0.0f, 0.0f);
        return m_rotate[std::min(k, m_rotate.size() - 1)];
    }

    // Interpolated components (RMath.h:681-715): the key verbatim when the mix
    // factor is exactly 0, else a*(1-t) + b*t; rotations renormalised.
    Vector translation(float time) const
    {
        if (m_time.empty()) return Vector(0.0f);
        float t;
        size_t i = bracket(time, t);
        return t == 0.0f ? m_translate[i] : m_translate[i] * (1.0f - t) + m_translate[i + 1] * t;
    }
    Vector scaling(float time) const
    {
        if (m_time.empty()) return Vector(1.0f);
        float t;
        size_t i = bracket(time, t);
        return t == 0.0f ? m_scale[i] : m_scale[i] * (1.0f - t) + m_scale[i + 1] * t;
    }
    Quaternion rotation(float time) const
    {
        if (m_time.empty()) return Quaternion(1.0f, 0.0f, 0.0f, 0.0f);
        float t;
        size_t i = bracket(time, t);
        return t == 0.0f ? m_rotate[i] : lerp(m_rotate[i], m_rotate[i + 1], t);
    }

    void setTranslationKey(size_t k, const Vector& v) { if (k < m_translate.size()) m_translate[k] = v; }
    void setScalingKey(size_t k, const Vector& v) { if (k < m_scale.size()) m_scale[k] = v; }
    void setRotationKey(size_t k, const Quaternion& q) { if (k < m_rotate.size()) m_rotate[k] = q; }
    void translateKey(size_t k, const Vector& v) { if (k < m_translate.size()) m_translate[k] += v; }
    void scaleKey(size_t k, const Vector& v) { if (k < m_scale.size()) m_scale[k] *= v; }
    void rotateKey(size_t k, const Quaternion& q) { if (k < m_rotate.size()) m_rotate[k] *= q; }

    void setTranslation(float time, const Vector& v) { m_translate[keyAt(time)] = v; }
    void setScaling(float time, const Vector& v) { m_scale[keyAt(time)] = v; }
    void setRotation(float time, const Quaternion& q) { m_rotate[keyAt(time)] = q; }
    void translate(float time, const Vector& v) { m_translate[keyAt(time)] += v; }
    void scale(float time, const Vector& v) { m_scale[keyAt(time)] *= v; }
    void rotate(float time, const Quaternion& q) { m_rotate[keyAt(time)] *= q; }

    // Normalise rotation keys (RMath.h:800-807).  Not idempotent in the last bit,
    // so a scene must be prepared exactly once, as in the reference GUI.
    void prepare()
    {
        for (size_t i = 0; i < m_rotate.size(); ++i)
            m_rotate[i].normalize();
    }

    Point toLocalPoint(float time, const Point& p) const { return ((~rotation(time)) * (p - translation(time))) / scaling(time); }
    Point fromLocalPoint(float time, const Point& p) const { return rotation(time) * (p * scaling(time)) + translation(time); }
    Vector toLocalVector(float time, const Vector& v) const { return ((~rotation(time)) * v) / scaling(time); }
    Vector fromLocalVector(float time, const Vector& v) const { return rotation(time) * (v * scaling(time)); }
    Vector toLocalNormal(float time, const Vector& n) const { return (~rotation(time)) * n; }
    Vector fromLocalNormal(float time, const Vector& n) const { return rotation(time) * n; }

    // Raw key storage, for flattening into RtSceneDesc
    const std::vector<float>& keyTimes() const { return m_time; }
    const std::vector<Vector>& scaleKeys() const { return m_scale; }
    const std::vector<Quaternion>& rotationKeys() const { return m_rotate; }
    const std::vector<Vector>& translationKeys() const { return m_translate; }

private:
    std::vector<float> m_time;
    std::vector<Vector> m_scale;
    std::vector<Quaternion> m_rotate;
    std::vector<Vector> m_translate;

    // Key just before `time` and the 0..1 mix towards the next one.  The search
    // and the pegging rules follow RMath.h:850-884 exactly (the device kernel
    // rt_xform.cuh implements the same walk).
    size_t bracket(float time, float& mix) const
    {
        size_t lo = 0, hi = m_time.size() - 1;
        if (m_time[hi] <= time) lo = hi;
        else if (m_time[lo] >= time) hi = lo;
        while (hi - lo > 0)
        {
            size_t mid = (lo + hi) / 2;
            if (time < m_time[mid]) hi = mid;
            else if (mid > lo) lo = mid;
            else break;
        }
        if (lo == m_time.size() - 1 || m_time[lo] >= time)
            mix = 0.0f;
        else
            mix = (time - m_time[lo]) / (m_time[lo + 1] - m_time[lo]);
        return lo;
    }

    // Find the key at `time`, creating it if needed (RMath.h:886-940): first key
    // is identity; past the end copies the last key; before the start copies the
    // first; between two keys inserts their interpolation.
    size_t keyAt(float time)
    {
        if (m_time.empty())
        {
            m_translate.push_back(Vector(0.0f));
            m_scale.push_back(Vector(1.0f));
            m_rotate.push_back(Quaternion(1.0f, 0.0f, 0.0f, 0.0f));
            m_time.push_back(time);
            return 0;
        }
        if (time > m_time.back())
        {
            m_translate.push_back(m_translate.back());
            m_scale.push_back(m_scale.back());
            m_rotate.push_back(m_rotate.back());
            m_time.push_back(time);
            return m_time.size() - 1;
        }
        if (time < m_time[0])
        {
            m_translate.insert(m_translate.begin(), m_translate.front());
            m_scale.insert(m_scale.begin(), m_scale.front());
            m_rotate.insert(m_rotate.begin(), m_rotate.front());
            m_time.insert(m_time.begin(), time);
            return 0;
        }
        float t;
        size_t i = bracket(time, t);
        if (t != 0.0f && t != 1.0f && i < m_time.size() - 1)
        {
            ++i;
            Vector trans = m_translate[i - 1] * (1.0f - t) + m_translate[i] * t;
            Vector scl = m_scale[i - 1] * (1.0f - t) + m_scale[i] * t;
            Quaternion rot = lerp(m_rotate[i - 1], m_rotate[i], t);
            m_translate.insert(m_translate.begin() + i, trans);
            m_scale.insert(m_scale.begin() + i, scl);
            m_rotate.insert(m_rotate.begin() + i, rot);
            m_time.insert(m_time.begin() + i, time);
        }
        return i;
    }
};

} // namespace Rayito

#endif // RAYITO_B200_MATH_HPP

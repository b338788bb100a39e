// Minimal stand-in for the Imath types in toucanRender's public signatures
// (V2i image size, V4f colours, Box2i fit boxes).  Imath is not available here.
#pragma once

#define IMATH_NAMESPACE Imath

namespace Imath {

template <class T>
struct Vec2 {
    T x = 0, y = 0;
    constexpr Vec2() = default;
    constexpr Vec2(T a, T b) : x(a), y(b) {}
    template <class S>
    constexpr Vec2(const Vec2<S>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    constexpr bool operator==(const Vec2& o) const { return x == o.x && y == o.y; }
    constexpr bool operator!=(const Vec2& o) const { return !(*this == o); }
};
template <class T>
struct Vec4 {
    T x = 0, y = 0, z = 0, w = 0;
    constexpr Vec4() = default;
    constexpr Vec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
};
template <class V>
struct Box {
    V min, max;
};
using V2i = Vec2<int>;
using V2d = Vec2<double>;
using V4f = Vec4<float>;
using Box2i = Box<V2i>;

} // namespace Imath

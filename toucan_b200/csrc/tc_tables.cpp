// tc_tables.cpp -- host-side constant tables for the kernels: colour-map knots
// (OIIO ImageBufAlgo::color_map, called from plugins/FilterPlugin.cpp:271-282) and the
// colour spaces of OpenColorIO's built-in config baked to  decode-curve -> 3x3 -> encode-curve
// (plugins/ColorPlugin.cpp:220-248).
#include <cmath>
#include <cstring>
#include <string>

namespace tc {

namespace {

// Knot tables of OIIO's color_map: 17 knots per matplotlib map / turbo, LINEAR light (generated, see the header of the
// include and tools/make_colormaps.py for the evidence and the precision of each table).
#include "tc_colormap_tables.inc"

struct MapEntry {
    const char* name;
    const float* knots;
    int n;
};
#define MAP(n, t) { n, t, (int)(sizeof(t) / sizeof(float) / 3) }
const MapEntry kMaps[] = {
    MAP("magma", kMap_magma), MAP("inferno", kMap_inferno), MAP("plasma", kMap_plasma), MAP("viridis", kMap_viridis),
    MAP("turbo", kMap_turbo), MAP("blue-red", kMap_bluered), MAP("red-blue", kMap_bluered), MAP("bluered", kMap_bluered),
    MAP("redblue", kMap_bluered), MAP("spectrum", kMap_spectrum), MAP("heat", kMap_heat),
};
#undef MAP

// ---- colour spaces -------------------------------------------------------------------
struct V3 {
    double v[3];
};
struct M3 {
    double m[3][3];
};
M3 mul(const M3& a, const M3& b)
{
    M3 r {};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
V3 mul(const M3& a, const V3& b)
{
    V3 r {};
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) r.v[i] += a.m[i][k] * b.v[k];
    return r;
}
M3 inverse(const M3& a)
{
    const double(*m)[3] = a.m;
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    M3 r;
    r.m[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det;
    r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det;
    r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det;
    r.m[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
    r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det;
    r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
    r.m[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det;
    r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det;
    r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
    return r;
}
V3 xyz(double x, double y) { return V3 { { x / y, 1.0, (1.0 - x - y) / y } }; }

struct Primaries {
    const char* id;
    double r[2], g[2], b[2], w[2];
};
const Primaries kPrims[] = {
    { "ap0", { 0.7347, 0.2653 }, { 0.0, 1.0 }, { 0.0001, -0.077 }, { 0.32168, 0.33767 } },
    { "ap1", { 0.713, 0.293 }, { 0.165, 0.830 }, { 0.128, 0.044 }, { 0.32168, 0.33767 } },
    { "rec709", { 0.64, 0.33 }, { 0.30, 0.60 }, { 0.15, 0.06 }, { 0.3127, 0.3290 } },
    { "p3d65", { 0.680, 0.320 }, { 0.265, 0.690 }, { 0.150, 0.060 }, { 0.3127, 0.3290 } },
    { "rec2020", { 0.708, 0.292 }, { 0.170, 0.797 }, { 0.131, 0.046 }, { 0.3127, 0.3290 } },
};
const Primaries* find_prims(const char* id)
{
    for (const auto& p : kPrims)
        if (!strcmp(p.id, id)) return &p;
    return nullptr;
}
// normalised primary matrix RGB -> XYZ
M3 npm(const Primaries& p)
{
    const V3 r = xyz(p.r[0], p.r[1]), g = xyz(p.g[0], p.g[1]), b = xyz(p.b[0], p.b[1]), w = xyz(p.w[0], p.w[1]);
    M3 P;
    for (int i = 0; i < 3; ++i) {
        P.m[i][0] = r.v[i];
        P.m[i][1] = g.v[i];
        P.m[i][2] = b.v[i];
    }
    const V3 S = mul(inverse(P), w);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P.m[i][j] *= S.v[j];
    return P;
}
// Bradford chromatic adaptation
M3 cat(const double sw[2], const double dw[2])
{
    const M3 B { { { 0.8951, 0.2664, -0.1614 }, { -0.7502, 1.7135, 0.0367 }, { 0.0389, -0.0685, 1.0296 } } };
    if (sw[0] == dw[0] && sw[1] == dw[1]) return M3 { { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } } };
    const V3 s = mul(B, xyz(sw[0], sw[1])), d = mul(B, xyz(dw[0], dw[1]));
    M3 D {};
    for (int i = 0; i < 3; ++i) D.m[i][i] = d.v[i] / s.v[i];
    return mul(inverse(B), mul(D, B));
}
M3 to_ap0(const Primaries& p)
{
    const Primaries& ap0 = kPrims[0];
    return mul(inverse(npm(ap0)), mul(cat(p.w, ap0.w), npm(p)));
}

enum { CURVE_LINEAR = 0, CURVE_GAMMA = 1, CURVE_SRGB = 2, CURVE_ACESCCT = 3, CURVE_ACESCC = 4 };
struct Space {
    const char* name;
    int curve;
    float param;
    const char* prims;
};
const Space kSpaces[] = {
    { "ACES2065-1", CURVE_LINEAR, 1.0f, "ap0" },
    { "ACEScg", CURVE_LINEAR, 1.0f, "ap1" },
    { "ACEScct", CURVE_ACESCCT, 1.0f, "ap1" },
    { "ACEScc", CURVE_ACESCC, 1.0f, "ap1" },
    { "Linear Rec.709 (sRGB)", CURVE_LINEAR, 1.0f, "rec709" },
    { "Linear P3-D65", CURVE_LINEAR, 1.0f, "p3d65" },
    { "Linear Rec.2020", CURVE_LINEAR, 1.0f, "rec2020" },
    { "Gamma 1.8 Rec.709 - Texture", CURVE_GAMMA, 1.8f, "rec709" },
    { "Gamma 2.2 Rec.709 - Texture", CURVE_GAMMA, 2.2f, "rec709" },
    { "Gamma 2.4 Rec.709 - Texture", CURVE_GAMMA, 2.4f, "rec709" },
    { "Gamma 2.2 AP1 - Texture", CURVE_GAMMA, 2.2f, "ap1" },
    { "sRGB - Texture", CURVE_SRGB, 1.0f, "rec709" },
    { "sRGB Encoded Rec.709 (sRGB)", CURVE_SRGB, 1.0f, "rec709" },
    { "sRGB Encoded AP1 - Texture", CURVE_SRGB, 1.0f, "ap1" },
    { "sRGB Encoded P3-D65 - Texture", CURVE_SRGB, 1.0f, "p3d65" },
};
const Space* find_space(const char* name)
{
    for (const auto& s : kSpaces)
        if (!strcmp(s.name, name)) return &s;
    return nullptr;
}

} // namespace

int color_map_knots(const char* name, const float** knots)
{
    for (const auto& m : kMaps)
        if (!strcmp(m.name, name)) {
            *knots = m.knots;
            return m.n;
        }
    return 0;
}

int color_space_known(const char* name) { return name && find_space(name) ? 1 : 0; }

int color_convert_setup(const char* from, const char* to, int* curve_in, float* p_in, float m[9], int* curve_out, float* p_out)
{
    const Space* f = find_space(from);
    const Space* t = find_space(to);
    if (!f || !t) return 0;
    *curve_in = f->curve;
    *p_in = f->param;
    *curve_out = t->curve;
    *p_out = t->param;
    M3 M { { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } } };
    if (strcmp(f->prims, t->prims) != 0) M = mul(inverse(to_ap0(*find_prims(t->prims))), to_ap0(*find_prims(f->prims)));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i * 3 + j] = (float)M.m[i][j];
    return 1;
}

} // namespace tc

extern "C" int tc_color_space_known(const char* name) { return tc::color_space_known(name); }

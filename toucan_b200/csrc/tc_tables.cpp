// tc_tables.cpp -- host-side constant tables for the kernels: colour-map knots
// (OIIO ImageBufAlgo::color_map, called from plugins/FilterPlugin.cpp:271-282) and the
// colour spaces of OpenColorIO's built-in config baked to  decode-curve -> 3x3 -> encode-curve
// (plugins/ColorPlugin.cpp:220-248).
#include <cmath>
#include <cstring>
#include <string>

namespace tc {

namespace {

// Knot tables as published with OIIO's color_map (6-decimal samples of the matplotlib maps).
const float kMagma[] = {
    0.001462f, 0.000466f, 0.013866f, 0.039608f, 0.031090f, 0.133515f, 0.113094f, 0.065492f, 0.276784f,
    0.211718f, 0.061992f, 0.418647f, 0.316654f, 0.071690f, 0.485380f, 0.414709f, 0.110431f, 0.504662f,
    0.512831f, 0.148179f, 0.507648f, 0.613617f, 0.181811f, 0.498536f, 0.716387f, 0.214982f, 0.475290f,
    0.816914f, 0.255895f, 0.436461f, 0.904281f, 0.319610f, 0.388137f, 0.960949f, 0.418323f, 0.359630f,
    0.986700f, 0.535582f, 0.382210f, 0.996096f, 0.653659f, 0.446213f, 0.996898f, 0.769591f, 0.534892f,
    0.992440f, 0.884330f, 0.640099f, 0.987053f, 0.991438f, 0.749504f,
};
const float kInferno[] = {
    0.001462f, 0.000466f, 0.013866f, 0.046915f, 0.030324f, 0.150164f, 0.142378f, 0.046242f, 0.308553f,
    0.258234f, 0.038571f, 0.406485f, 0.366529f, 0.071579f, 0.431994f, 0.472328f, 0.110547f, 0.428334f,
    0.578304f, 0.148039f, 0.404411f, 0.682656f, 0.189501f, 0.360757f, 0.780517f, 0.243327f, 0.299523f,
    0.865006f, 0.316822f, 0.226055f, 0.929644f, 0.411479f, 0.145367f, 0.970919f, 0.522853f, 0.058367f,
    0.987622f, 0.645320f, 0.039886f, 0.978806f, 0.774545f, 0.176037f, 0.950018f, 0.903409f, 0.380271f,
    0.988362f, 0.998364f, 0.644924f,
};
const float kPlasma[] = {
    0.050383f, 0.029803f, 0.527975f, 0.186213f, 0.018803f, 0.587228f, 0.287076f, 0.010855f, 0.627295f,
    0.381047f, 0.001814f, 0.653068f, 0.471457f, 0.005678f, 0.659897f, 0.557243f, 0.047331f, 0.643443f,
    0.636008f, 0.112092f, 0.605205f, 0.706178f, 0.178437f, 0.553657f, 0.768090f, 0.244817f, 0.498465f,
    0.823132f, 0.311261f, 0.444806f, 0.872303f, 0.378774f, 0.393355f, 0.915471f, 0.448807f, 0.342890f,
    0.951344f, 0.522850f, 0.292275f, 0.977856f, 0.602051f, 0.241387f, 0.992541f, 0.687030f, 0.192170f,
    0.992505f, 0.777967f, 0.152855f, 0.940015f, 0.975158f, 0.131326f,
};
const float kViridis[] = {
    0.267004f, 0.004874f, 0.329415f, 0.282656f, 0.100196f, 0.422160f, 0.277134f, 0.185228f, 0.489898f,
    0.253935f, 0.265254f, 0.529983f, 0.221989f, 0.339161f, 0.548752f, 0.190631f, 0.407061f, 0.556089f,
    0.163625f, 0.471133f, 0.558148f, 0.139147f, 0.533812f, 0.555298f, 0.120565f, 0.596422f, 0.543611f,
    0.134692f, 0.658636f, 0.517649f, 0.208030f, 0.718701f, 0.472873f, 0.327796f, 0.773980f, 0.406640f,
    0.477504f, 0.821444f, 0.318195f, 0.647257f, 0.858400f, 0.209861f, 0.824940f, 0.884720f, 0.106217f,
    0.993248f, 0.906157f, 0.143936f,
};
// turbo at 8-bit precision (the 6-decimal table is not restated; documented in DESIGN.md)
#define U8(v) ((v) / 255.f)
const float kTurbo[] = {
    U8(48), U8(18), U8(59), U8(65), U8(69), U8(171), U8(70), U8(117), U8(237), U8(57), U8(162), U8(252),
    U8(27), U8(207), U8(212), U8(36), U8(236), U8(166), U8(97), U8(252), U8(108), U8(164), U8(252), U8(59),
    U8(209), U8(232), U8(52), U8(243), U8(198), U8(58), U8(254), U8(155), U8(45), U8(243), U8(99), U8(21),
    U8(217), U8(56), U8(6), U8(177), U8(25), U8(1), U8(122), U8(4), U8(2),
};
#undef U8
const float kBlueRed[] = { 0.0f, 0.0f, 1.0f, 1.0f, 0.0f, 0.0f };
const float kSpectrum[] = { 0, 0, 0.05f, 0, 0, 0.75f, 0, 0.5f, 0, 0.5f, 0.5f, 0, 1, 0, 0 };
const float kHeat[] = { 0, 0, 0, 0.05f, 0, 0, 0.25f, 0, 0, 0.75f, 0.75f, 0, 1, 1, 1 };

struct MapEntry {
    const char* name;
    const float* knots;
    int n;
};
#define MAP(n, t) { n, t, (int)(sizeof(t) / sizeof(float) / 3) }
const MapEntry kMaps[] = {
    MAP("magma", kMagma), MAP("inferno", kInferno), MAP("plasma", kPlasma), MAP("viridis", kViridis),
    MAP("turbo", kTurbo), MAP("blue-red", kBlueRed), MAP("red-blue", kBlueRed), MAP("bluered", kBlueRed),
    MAP("redblue", kBlueRed), MAP("spectrum", kSpectrum), MAP("heat", kHeat),
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

"""GPU parity: every C-ABI kernel against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact for Fill / Checkers / Flip / Flop / Crop / Invert /
Wipes / seeded Noise (and everything else that is pure select/copy); <= 1e-5 abs (float32)
or +-1 LSB (u8 / u16 / half storage) for filtering, resampling and colour ops.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

NP_TYPES = [np.uint8, np.uint16, np.float16, np.float32]
SIZES = [(333, 217), (64, 48), (1, 1), (37, 5)]


def rand_img(w, h, dt, seed, alpha="frac"):
    rng = np.random.default_rng(seed)
    f = rng.random((h, w, 4), dtype=np.float32)
    if alpha == "one":
        f[..., 3] = 1.0
    elif alpha == "mixed":
        f[..., 3] = np.where(rng.random((h, w)) < 0.3, 1.0, np.where(rng.random((h, w)) < 0.3, 0.0, f[..., 3]))
    if dt == np.uint8:
        return np.ascontiguousarray((f * 255).astype(np.uint8))
    if dt == np.uint16:
        return np.ascontiguousarray((f * 65535).astype(np.uint16))
    return np.ascontiguousarray(f.astype(dt))


def to_dev(a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def to_host(t):
    if isinstance(t, np.ndarray):
        return t
    if t.dtype == torch.uint16:
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def assert_exact(got, want, what=""):
    got = to_host(got)
    assert got.dtype == want.dtype and got.shape == want.shape, (what, got.dtype, want.dtype, got.shape, want.shape)
    if got.dtype.kind == "f":
        same = (got.view(np.uint16 if got.dtype == np.float16 else np.uint32) ==
                want.view(np.uint16 if got.dtype == np.float16 else np.uint32)) | ((got == want))
        bad = ~same
    else:
        bad = got != want
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} values differ; first at {np.argwhere(bad)[0]}: got {got[bad][0]} want {want[bad][0]}"


def assert_close(got, want, what="", tol=1e-5, lsb=1, rtol=0.0):
    got = to_host(got)
    assert got.dtype == want.dtype and got.shape == want.shape, (what, got.dtype, want.dtype, got.shape, want.shape)
    if got.dtype == np.float32:
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        assert np.isfinite(got).all() or not np.isfinite(want).all(), what
        d = d - rtol * np.abs(want.astype(np.float64))
        m = float(np.nanmax(d)) if d.size else 0.0
        assert m <= tol, f"{what}: max abs diff {m:.3e} > {tol}"
    elif got.dtype == np.float16:
        # +-1 LSB of half storage: neighbouring representable values
        gi = got.view(np.int16).astype(np.int32)
        wi = want.view(np.int16).astype(np.int32)
        gi = np.where(gi < 0, -(gi & 0x7FFF), gi)
        wi = np.where(wi < 0, -(wi & 0x7FFF), wi)
        ulp = np.abs(gi - wi)
        # near zero a 1e-6 difference spans several (sub)normal half steps: accept the float bar too
        bad = (ulp > lsb) & (np.abs(got.astype(np.float64) - want.astype(np.float64)) > tol)
        assert not bad.any(), f"{what}: max half-ULP diff {int(ulp[bad].max())} at {np.argwhere(bad)[0]}"
    else:
        m = int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max()) if got.size else 0
        assert m <= lsb, f"{what}: max LSB diff {m}"


@pytest.fixture(scope="module")
def T(cuda):
    from toucan_b200 import ops

    return ops


C1 = [0.25, 0.5, 0.75, 1.0]
C2 = [1.0, 0.1, 0.333, 0.6]


@pytest.mark.parametrize("size", [(1280, 720), (333, 217), (1, 1)])
def test_generators_bit_exact(T, size):
    w, h = size
    assert_exact(T.fill(w, h, C1), O.fill(w, h, C1), "fill")
    assert_exact(T.fill(w, h, [0.0, 0.0, 1.0, 1.0]), O.fill(w, h, [0.0, 0.0, 1.0, 1.0]), "fill")
    for cs in [(100, 100), (200, 200), (300, 300), (7, 3)]:
        assert_exact(T.checkers(w, h, cs[0], cs[1], C1, C2), O.checkers(w, h, cs[0], cs[1], C1, C2), f"checkers {cs}")
    for vertical in (True, False):
        assert_exact(T.gradient(w, h, C1, C2, vertical), O.gradient(w, h, C1, C2, vertical), f"gradient v={vertical}")
        assert_exact(T.gradient(w, h, [1, 0, 0, 1], [0, 0, 1, 1], vertical), O.gradient(w, h, [1, 0, 0, 1], [0, 0, 1, 1], vertical), "gradient")


@pytest.mark.parametrize("dt", NP_TYPES)
def test_generators_other_types(T, dt):
    t = O._DT[np.dtype(dt)]
    w, h = 129, 67
    assert_exact(T.fill(w, h, C2, t), O.fill(w, h, C2, t), "fill")
    assert_exact(T.checkers(w, h, 10, 9, C1, C2, t), O.checkers(w, h, 10, 9, C1, C2, t), "checkers")
    assert_exact(T.gradient(w, h, C1, C2, True, t), O.gradient(w, h, C1, C2, True, t), "gradient v")
    assert_exact(T.gradient(w, h, C1, C2, False, t), O.gradient(w, h, C1, C2, False, t), "gradient h")


@pytest.mark.parametrize("ntype,a,b,mono,seed", [("gaussian", 1.0, 1.0, False, 0), ("gaussian", 0.5, 0.1, True, 7),
                                                 ("uniform", 0.1, 0.9, False, 3), ("white", 0.0, 1.0, True, 11),
                                                 ("salt", 1.0, 0.05, False, 5), ("bogus", 1.0, 1.0, False, 0)])
def test_noise_bit_exact(T, ntype, a, b, mono, seed):
    for (w, h) in [(1280, 720), (97, 33)]:
        assert_exact(T.noise(w, h, ntype, a, b, mono, seed), O.noise(w, h, ntype, a, b, mono, seed), f"noise {ntype} u8")
    t = O.F32
    assert_exact(T.noise(333, 217, ntype, a, b, mono, seed, t), O.noise(333, 217, ntype, a, b, mono, seed, t), f"noise {ntype} f32")


def test_noise_gaussian_bit_exact_over_a_large_sample(T):
    # 2 x 1920 x 1080 x 4 = 16.6 M float samples: the kernel's Newton / fused-multiply-add evaluation of sqrt(-2 ln r / r) with its
    # exact-chain decision near float rounding boundaries (tests/cpp/noise_fastpath_check.c) against the oracle's IEEE double chain
    for seed in (1, 2026):
        assert_exact(T.noise(1920, 1080, "gaussian", 0.5, 0.25, False, seed, O.F32), O.noise(1920, 1080, "gaussian", 0.5, 0.25, False, seed, O.F32),
                     f"noise gaussian f32 seed {seed}")


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("size", SIZES)
def test_exact_unary(T, dt, size):
    w, h = size
    src = rand_img(w, h, dt, 1)
    d = to_dev(src)
    assert_exact(T.invert(d), O.invert(src), "invert")
    assert_exact(T.flip(d), O.flip(src), "flip")
    assert_exact(T.flop(d), O.flop(src), "flop")
    assert_exact(T.flip(T.flip(d)), src, "flip∘flip")
    pos, csz = (w // 3, h // 4), (max(1, w // 2), max(1, h // 2))
    assert_exact(T.crop(d, pos, csz), O.crop(src, pos, csz), "crop")
    assert_exact(T.crop(d, (w // 2, h // 2), (w, h)), O.crop(src, (w // 2, h // 2), (w, h)), "crop past the edge")
    for t in range(4):
        assert_exact(T.convert(d, t), O.convert(src, t), f"convert ->{t}")


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("size", SIZES[:2])
def test_pointwise_filters(T, dt, size):
    w, h = size
    src = rand_img(w, h, dt, 2, alpha="mixed")
    d = to_dev(src)
    assert_close(T.saturate(d, 0.0), O.saturate(src, 0.0), "saturate 0")
    assert_close(T.saturate(d, 1.7), O.saturate(src, 1.7), "saturate 1.7")
    assert_close(T.pow_(d, 2.0), O.pow_(src, 2.0), "pow 2")
    assert_close(T.pow_(d, 3.0), O.pow_(src, 3.0), "pow 3")
    assert_close(T.pow_(d, 0.4545), O.pow_(src, 0.4545), "pow .45")
    for name in ["plasma", "inferno", "magma", "viridis", "turbo", "blue-red", "spectrum", "heat", "nonesuch"]:
        assert_close(T.color_map(d, name), O.color_map(src, name), f"color_map {name}")
    assert_close(T.premult(d), O.premult(src), "premult")
    assert_close(T.unpremult(d), O.unpremult(src), "unpremult")
    assert_close(T.colorconvert(d, "Linear Rec.709 (sRGB)", "ACEScct"), O.colorconvert(src, "Linear Rec.709 (sRGB)", "ACEScct"), "cc1")
    assert_close(T.colorconvert(d, "Gamma 2.4 Rec.709 - Texture", "Linear P3-D65"),
                 O.colorconvert(src, "Gamma 2.4 Rec.709 - Texture", "Linear P3-D65"), "cc2")
    assert_close(T.colorconvert(d, "sRGB - Texture", "ACEScg"), O.colorconvert(src, "sRGB - Texture", "ACEScg"), "cc3")
    assert_close(T.colorconvert(d, "ACEScc", "Linear Rec.2020"), O.colorconvert(src, "ACEScc", "Linear Rec.2020"), "cc4", rtol=1e-6)  # ACEScc decode reaches ~220: relative bar
    assert_close(T.colorconvert(d, "nope", "ACEScg"), O.colorconvert(src, "nope", "ACEScg"), "cc unknown")
    # OIIO's unpremult flag (toucan's "premult" parameter): divide by alpha, convert, multiply back
    assert_close(T.colorconvert(d, "sRGB - Texture", "ACEScg", unpremult=True), O.colorconvert(src, "sRGB - Texture", "ACEScg", unpremult=True),
                 "cc unpremult", rtol=2e-6)


@pytest.mark.parametrize("value", [2.2, 0.4545, -1.5, 4.0, 5.0, -3.0, 0.0, 1.0, 7.25])
def test_pow_special_values(T, value):
    """toucan:Pow is powf per channel (FilterPlugin.cpp:420-441): negative bases (NaN unless the exponent is an integer,
    then signed by parity), zeros, denormals, infinities, NaN, and magnitudes from 1e-30 to 1e30 -- relative bar 1e-6
    (exponents of a few units: the error of the split-log2 evaluation grows with |value|)."""
    rng = np.random.default_rng(11)
    mags = np.float32(10.0) ** rng.uniform(-30, 30, size=(64, 64, 4)).astype(np.float32)
    src = (mags * np.where(rng.random((64, 64, 4)) < 0.3, -1, 1)).astype(np.float32)
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 1.17549435e-38, 3.4028235e38, 0.5, 2.0],
                       dtype=np.float32)
    src.reshape(-1)[:special.size] = special
    with np.errstate(all="ignore"):
        want = O.pow_(src, value)
    got = to_host(T.pow_(to_dev(src), value))
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern"
    inf = np.isinf(want)
    big = np.abs(want) > 1e37  # within a few ulp of overflow either side may round to infinity
    assert np.array_equal(got[inf & ~big], want[inf & ~big]) and np.array_equal(np.isinf(got) & ~big, inf & ~big), "infinities"
    fin = np.isfinite(want) & np.isfinite(got) & ~big
    tiny = np.abs(want) < 1e-37  # denormal results: absolute bar
    d = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert (d[fin & tiny] <= 1e-37).all()
    rel = d[fin & ~tiny] / np.abs(want[fin & ~tiny].astype(np.float64))
    assert rel.max() <= 1e-6, f"pow {value}: max relative error {rel.max():.3e}"
    assert np.array_equal(np.signbit(got[fin & ~tiny]), np.signbit(want[fin & ~tiny])), "sign"


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("radius", [10.0, 20.0, 3.0, 1.0, 7.5, 13.0, 25.0])  # tile templates R = 4, 9, 2, 2, 4, 6, 12
def test_blur_unsharp(T, dt, radius):
    # 332 x 217: rows are 16-byte multiples for every type -> the TMA kernel (k_gauss_tma), partial tiles on both axes;
    # 333 x 217: odd 8/16-bit row pitches -> the cp-free typed-load kernel (k_gauss_tile); 40 x 30: smaller than a tile
    for (w, h) in [(332, 217), (333, 217), (40, 30)]:
        src = rand_img(w, h, dt, 3)
        d = to_dev(src)
        assert_close(T.blur(d, radius), O.blur(src, radius), f"blur {radius}")
        assert_close(T.unsharp_mask(d, "gaussian", radius, 1.0, 0.0), O.unsharp_mask(src, "gaussian", radius, 1.0, 0.0), "unsharp")
        # the threshold is a discontinuity (|src - blurry| < threshold -> no change): values whose difference sits within
        # 1e-4 of it may fall on either side for a 1e-7 difference in the blur and are left out of the comparison
        srcf = src.astype(np.float32) / (255.0 if dt == np.uint8 else 65535.0 if dt == np.uint16 else 1.0)
        near = np.abs(np.abs(srcf - O.blur(np.ascontiguousarray(srcf), radius)) - 0.05) < 1e-4
        got = to_host(T.unsharp_mask(d, "gaussian", radius, 0.7, 0.05))
        want = O.unsharp_mask(src, "gaussian", radius, 0.7, 0.05)
        assert near.mean() < 0.01
        assert_close(np.where(near, want, got), want, "unsharp thr", lsb=1)


def test_blur_constant_image(T):
    # known answer: blurring a constant image returns the constant (to rounding)
    src = np.full((50, 70, 4), 0.375, np.float32)
    got = to_host(T.blur(to_dev(src), 10.0))
    assert np.abs(got - 0.375).max() < 1e-6


@pytest.mark.parametrize("dt", NP_TYPES)
def test_resize(T, dt):
    src = rand_img(160, 90, dt, 4)
    d = to_dev(src)
    # the march kernel with 1 .. 4 staged column spans ((48, 27): 125 staged columns); (33, 77), (20, 12): a footprint beyond its
    # tile -> two-pass path with the staged row pass; (5, 3): beyond that -> gathers
    for size in [(320, 180), (80, 45), (160, 90), (200, 50), (48, 27), (33, 77), (20, 12), (5, 3)]:
        assert_close(T.resize(d, size), O.resize(src, size), f"resize {size}")
    # tiny and one-pixel-wide sources, a tall frame (several chunks of bands per strip), a large enlargement (hundreds of rows per band)
    for (w, h), size in [((1, 1), (7, 5)), ((3, 2), (64, 64)), ((2, 300), (3, 1000)), ((31, 1000), (17, 333)), ((40, 24), (400, 250))]:
        s2 = rand_img(w, h, dt, 9)
        assert_close(T.resize(to_dev(s2), size), O.resize(s2, size), f"resize {(w, h)} -> {size}")
    assert_close(T.resize(d, (100, 60), "lanczos3", 0.0), O.resize(src, (100, 60), "lanczos3", 0.0), "resize lanczos3")
    assert_close(T.resize(d, (300, 200), "blackman-harris", 0.0), O.resize(src, (300, 200), "blackman-harris", 0.0), "resize bh")


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("angle", [45.0, 0.0, 90.0, -30.0, 360.0, 181.5])
def test_rotate(T, dt, angle):
    src = rand_img(160, 90, dt, 5)
    assert_close(T.rotate(to_dev(src), angle), O.rotate(src, angle), f"rotate {angle}")


@pytest.mark.parametrize("size", [(1, 1), (37, 5), (5, 37), (64, 48), (333, 217)])
def test_rotate_small_and_odd_frames(T, size):
    # frames smaller than the staged box (the TMA copy reads mostly outside the frame: zero fill = the black surround)
    for dt in (np.float32, np.uint8):
        src = rand_img(size[0], size[1], dt, 11)
        for angle in (45.0, -17.0):
            assert_close(T.rotate(to_dev(src), angle), O.rotate(src, angle), f"rotate {size} {angle}")


@pytest.mark.parametrize("dta", NP_TYPES)
@pytest.mark.parametrize("dtb", NP_TYPES)
@pytest.mark.parametrize("size", [(333, 217), (64, 48)])  # odd pixel count: scalar path; multiple of 4: the 128-bit groups, mixed types included
def test_transitions_and_over(T, dta, dtb, size):
    w, h = size
    a, b = rand_img(w, h, dta, 6), rand_img(w, h, dtb, 7)
    da, db = to_dev(a), to_dev(b)
    exact_ok = True
    for v in [0.0, 1.0 / 6.0, 0.5, 5.0 / 6.0, 1.0]:
        assert_exact(T.dissolve(da, db, v), O.dissolve(a, b, v), f"dissolve {v}")
        for vertical in (False, True):
            assert_exact(T.wipe(da, db, v, vertical), O.wipe(a, b, v, vertical), f"wipe {v} {vertical}")
    assert_exact(T.over(da, db, premult_fg=True), O.over(O.premult(a), b), "premult+over")
    assert_exact(T.over(da, db, premult_fg=False), O.over(a, b), "over")
    assert exact_ok


def test_wipe_wide_frame(T):
    # the 200 px ramp needs a frame wider than the fixtures' test sizes
    a, b = rand_img(1280, 720, np.uint8, 8, "one"), rand_img(1280, 720, np.uint8, 9)
    da, db = to_dev(a), to_dev(b)
    for v in [0.0, 0.25, 0.5, 0.75, 0.999]:
        assert_exact(T.wipe(da, db, v, False), O.wipe(a, b, v, False), "hwipe")
        assert_exact(T.wipe(da, db, v, True), O.wipe(a, b, v, True), "vwipe")


def test_dissolve_known_answers(T):
    a, b = rand_img(64, 48, np.float32, 10), rand_img(64, 48, np.float32, 11)
    assert_exact(T.dissolve(to_dev(a), to_dev(b), 0.0), a, "dissolve(0) = from")
    assert_exact(T.dissolve(to_dev(a), to_dev(b), 1.0), b, "dissolve(1) = to")


def test_paste_and_comp_resize(T):
    # CompNode's fit-resize path: resize fg into bg's box, paste on transparent, over
    fg = rand_img(64, 36, np.uint8, 12)
    bg = rand_img(128, 72, np.uint8, 13, "one")
    want = O.comp(fg, bg, True)
    dfg = T.premult(to_dev(fg))
    box = O.fit((128, 72), (64, 36))
    r = T.resize(dfg, (box[2], box[3]))
    p = T.paste_on_black(r, (128, 72), box[0], box[1], O.U8)
    assert_close(T.over(p, to_dev(bg)), want, "comp with fit-resize")


@pytest.mark.parametrize("n_layers,trans", [(1, False), (3, True), (10, True)])
def test_stack_fused_matches_unfused_and_oracle(T, n_layers, trans):
    w, h = 200, 120
    rng = np.random.default_rng(20)
    layers_np, layers_dev = [], []
    for l in range(n_layers):
        a = rand_img(w, h, np.float32, 100 + l)
        b = rand_img(w, h, np.float32, 200 + l) if (trans and l % 2 == 0) else None
        v = float(rng.random())
        layers_np.append((a, b, 10.0, 10.0, v))
        layers_dev.append((to_dev(a), None if b is None else to_dev(b), 10.0, 10.0, v))
    # oracle: the unfused node chain
    acc = O.fill(w, h, [0, 0, 0, 1])
    for a, b, ra, rb, v in layers_np:
        x = O.blur(a, ra)
        if b is not None:
            x = O.dissolve(x, O.blur(b, rb), v)
        acc = O.comp(x, acc, True)
    got = T.stack_f32(w, h, [0, 0, 0, 1], layers_dev)
    assert_close(got, acc, "fused stack vs oracle chain")
    # unfused CUDA chain
    acc_d = T.fill(w, h, [0, 0, 0, 1])
    for a, b, ra, rb, v in layers_dev:
        x = T.blur(a, ra)
        if b is not None:
            x = T.dissolve(x, T.blur(b, rb), v)
        acc_d = T.over(x, acc_d, premult_fg=True)
    assert_close(got, to_host(acc_d), "fused stack vs unfused kernels")


@pytest.mark.parametrize("radius", [30.0, 49.0])
def test_blur_wide_kernel_fallback(T, radius):
    # more than 12 live taps per side: the generic wide-kernel path
    src = rand_img(150, 97, np.float32, 21)
    assert_close(T.blur(to_dev(src), radius), O.blur(src, radius), f"blur {radius}")
    assert_close(T.unsharp_mask(to_dev(src), "gaussian", radius, 0.5, 0.0), O.unsharp_mask(src, "gaussian", radius, 0.5, 0.0), "unsharp wide")


@pytest.mark.parametrize("radius", [3.0, 7.0, 10.0, 14.0, 20.0, 26.0])
def test_stack_other_radii_and_mixed_layers(T, radius):
    # template instantiations R = 2, 4, 6, 9, 12; blurred, unblurred and dissolve layers in one stack;
    # odd frame size (partial tiles on both axes)
    w, h = 203, 77
    a = [rand_img(w, h, np.float32, 300 + i) for i in range(6)]
    d = [to_dev(x) for x in a]
    layers_np = [(a[0], None, radius, 0.0, 0.0), (a[1], a[2], radius, radius, 0.3), (a[3], None, 0.0, 0.0, 0.0),
                 (a[4], a[5], 0.0, 0.0, 0.75)]
    layers_dev = [(d[0], None, radius, 0.0, 0.0), (d[1], d[2], radius, radius, 0.3), (d[3], None, 0.0, 0.0, 0.0),
                  (d[4], d[5], 0.0, 0.0, 0.75)]
    acc = O.fill(w, h, [0.1, 0.2, 0.3, 1.0], O.F32)
    for sa, sb, ra, rb, v in layers_np:
        x = O.blur(sa, ra) if ra > 0 else sa
        if sb is not None:
            x = O.dissolve(x, O.blur(sb, rb) if rb > 0 else sb, v)
        acc = O.comp(x, acc, True)
    got = T.stack_f32(w, h, [0.1, 0.2, 0.3, 1.0], layers_dev)
    assert_close(got, acc, f"stack radius {radius}")


def test_stack_rejects_what_it_cannot_fuse(T):
    from toucan_b200 import capi

    a, b = to_dev(rand_img(64, 48, np.float32, 1)), to_dev(rand_img(64, 48, np.float32, 2))
    with pytest.raises(capi.TcError):  # different blurs on the two dissolve inputs
        T.stack_f32(64, 48, [0, 0, 0, 1], [(a, b, 10.0, 4.0, 0.5)])
    with pytest.raises(capi.TcError):  # radius beyond the fused kernel's range
        T.stack_f32(64, 48, [0, 0, 0, 1], [(a, None, 60.0, 0.0, 0.0)])


# ---- draw (plugins/DrawPlugin.cpp) -------------------------------------------------
@pytest.mark.parametrize("dt", NP_TYPES)
def test_draw_box_bit_exact(dt):
    from toucan_b200 import ops

    w, h = 97, 61
    src = rand_img(w, h, dt, 40)
    cases = [((10, 8), (60, 40), (0.8, 0.0, 0.3, 1.0), True),       # opaque fill
             ((10, 8), (60, 40), (0.4, 0.1, 0.3, 0.5), True),       # "over" with alpha < 1
             ((-20, -5), (30, 200), (0.2, 0.9, 0.3, 0.25), True),   # clipped on three sides
             ((60, 40), (10, 8), (1.0, 1.0, 1.0, 1.0), True),       # corners out of order: empty ROI
             ((5, 5), (90, 55), (0.9, 0.5, 0.1, 0.7), False),       # outline = four skip-first lines
             ((-3, 20), (120, 70), (0.9, 0.5, 0.1, 1.0), False),    # outline partly outside
             ((33, 12), (33, 12), (0.1, 0.2, 0.3, 0.4), True),      # degenerate: one point
             ((200, 200), (200, 200), (0.1, 0.2, 0.3, 0.4), False)]  # a point outside the image
    for p1, p2, color, fill in cases:
        assert_exact(ops.draw_box(to_dev(src), p1, p2, color, fill), O.draw_box(src, p1, p2, color, fill), f"box {p1} {p2} {color} {fill}")


@pytest.mark.parametrize("dt", [np.uint8, np.float32])
def test_draw_line_matches_the_sequential_bresenham_walk(dt):
    """Every step of the line is an independent thread (closed-form minor-axis advance); the oracle walks OIIO's
    sequential Bresenham loop.  All octants, degenerate and clipped lines, with and without the first point."""
    from toucan_b200 import ops

    w, h = 83, 59
    src = rand_img(w, h, dt, 41)
    rng = np.random.default_rng(42)
    lines = [((0, 0), (82, 58)), ((82, 58), (0, 0)), ((0, 58), (82, 0)), ((10, 5), (10, 50)), ((70, 30), (3, 30)), ((4, 4), (4, 4)),
             ((-30, 10), (120, 44)), ((40, -20), (45, 90)), ((0, 0), (82, 1)), ((0, 0), (1, 58)), ((5, 7), (6, 8)), ((20, 20), (21, 20))]
    lines += [(tuple(rng.integers(-20, 110, 2).tolist()), tuple(rng.integers(-20, 90, 2).tolist())) for _ in range(60)]
    for i, (p1, p2) in enumerate(lines):
        color = (0.9, 0.5, 0.1, 1.0) if i % 2 else (0.3, 0.2, 0.1, 0.6)
        for skip in (False, True):
            assert_exact(ops.draw_line(to_dev(src), p1, p2, color, skip), O.draw_line(src, p1, p2, color, skip), f"line {p1} {p2} skip={skip}")


def test_draw_long_lines_at_8k():
    """BASELINE size: a diagonal across a 7680x4320 frame (7680 threads), checked through its drawn pixels only."""
    from toucan_b200 import ops

    w, h = 7680, 4320
    src = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    out = ops.draw_line(src, (0, 0), (w - 1, h - 1), (1.0, 1.0, 1.0, 1.0))
    ys, xs = torch.nonzero(out[..., 0], as_tuple=True)
    assert xs.numel() == w and torch.equal(torch.sort(xs).values, torch.arange(w, device="cuda"))  # one pixel per column
    want = O.draw_line(np.zeros((64, 114, 4), np.uint8), (0, 0), (113, 63), (1.0, 1.0, 1.0, 1.0))  # same slope class, small
    small = ops.draw_line(torch.zeros((64, 114, 4), dtype=torch.uint8, device="cuda"), (0, 0), (113, 63), (1.0, 1.0, 1.0, 1.0))
    assert np.array_equal(small.cpu().numpy(), want)
    assert int(ys.max()) == h - 1 and int(ys.min()) == 0


# ---- chain of pointwise effects in one pass (cross-node fusion) ---------------------------------------
CHAINS = [
    [("color_map", "plasma"), ("invert",), ("pow", 2.0), ("saturate", 0.0)],                 # SURVEY 8d config 2, chained
    [("unpremult",), ("color_convert", "Linear Rec.709 (sRGB)", "ACEScct", False), ("premult",)],
    [("pow", 2.2), ("color_convert", "sRGB - Texture", "no such space", False), ("invert",)],   # failed convert: zeros, then invert
    [("saturate", 1.5), ("color_map", "no such map"), ("invert",), ("pow", 3.0), ("premult",), ("unpremult",), ("invert",), ("saturate", 0.25)],
    # data/MultipleEffects.otio's second clip through the whole graph: Flop, ColorMap (clip), Invert (track), CompNode over the
    # Fill frame (premult + over), Pow 3 (stack) -- five graph nodes, one launch
    [("flop",), ("color_map", "plasma"), ("invert",), ("over_fill", (0.0, 0.0, 0.0, 1.0), True), ("pow", 3.0)],
    [("invert",), ("flip",), ("saturate", 0.3), ("flop",), ("over_fill", (0.25, 0.5, 0.125, 0.75), True)],
    [("flip",), ("flip",), ("flop",), ("over_fill", (1.0, 0.0, 0.0, 0.5), False)],
]


def _single(ops_mod, t, st):
    return {"invert": lambda: ops_mod.invert(t), "pow": lambda: ops_mod.pow_(t, st[1]), "saturate": lambda: ops_mod.saturate(t, st[1]),
            "color_map": lambda: ops_mod.color_map(t, st[1]), "premult": lambda: ops_mod.premult(t), "unpremult": lambda: ops_mod.unpremult(t),
            "color_convert": lambda: ops_mod.colorconvert(t, st[1], st[2], unpremult=st[3]),
            "flip": lambda: ops_mod.flip(t), "flop": lambda: ops_mod.flop(t),
            "over_fill": lambda: _over_fill(ops_mod, t, st[1], st[2])}[st[0]]()


def _over_fill(ops_mod, t, color, premult):
    """CompNode over the Fill generator's UINT8 frame, node by node: fill, then premult + over."""
    h, w = t.shape[0], t.shape[1]
    if ops_mod is O:
        return O.comp(t, O.fill(w, h, list(color)), premult)
    return ops_mod.over(t, ops_mod.fill(w, h, list(color)), premult_fg=premult)


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("chain", range(len(CHAINS)))
def test_pointwise_chain_equals_the_nodes_one_by_one(dt, chain):
    """Bit-identical to the per-node kernels run in sequence (the storage round trip of every node boundary is
    replayed in registers), for every storage type; and within tolerance of the oracle's node-by-node evaluation."""
    from toucan_b200 import ops

    steps = CHAINS[chain]
    src = rand_img(171, 93, dt, 50 + chain, alpha="mixed")
    got = ops.pointwise_chain(to_dev(src), steps)
    seq = to_dev(src)
    ref = src
    for st in steps:
        seq = _single(ops, seq, st)
        ref = _single(O, ref, st)
    assert_exact(got, to_host(seq), f"chain {chain} vs per-node kernels")
    assert_close(got, ref, f"chain {chain} vs oracle", tol=2e-5, lsb=2)


CONV_CHAINS = [
    # data/MultipleEffects.otio's first clip below its Flip: Blur (clip), Invert (track), CompNode over the Fill frame, Pow 3 (stack):
    # simple followers -> the Blur kernel's epilogue, one launch
    [("blur", 20.0), ("invert",), ("over_fill", (0.0, 0.0, 0.0, 1.0), True), ("pow", 3.0)],
    [("blur", 10.0), ("over_fill", (0.2, 0.4, 0.6, 1.0), True)],
    # SURVEY 8d config 2's tail: UnsharpMask, then CompNode over the Fill frame
    [("unsharp", "gaussian", 10.0, 1.0, 0.0), ("over_fill", (0.0, 0.0, 0.0, 1.0), True)],
    [("unsharp", "gaussian", 3.0, 0.7, 0.05), ("premult",), ("pow", 2.0), ("invert",)],
    # followers the epilogue does not cover (Saturate, ColorMap, Pow 2.2): convolution into a temporary, then the rest as one launch
    [("blur", 7.0), ("saturate", 0.5), ("color_map", "plasma"), ("invert",)],
    [("blur", 13.0), ("pow", 2.2), ("over_fill", (0.0, 0.0, 0.0, 1.0), True)],
]


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("chain", range(len(CONV_CHAINS)))
@pytest.mark.parametrize("size", [(332, 217), (333, 217)])  # TMA kernel / typed-load kernel (odd 8/16-bit row pitch: not fusable into the kernel)
def test_chain_with_a_convolution_head_equals_the_nodes_one_by_one(dt, chain, size):
    from toucan_b200 import capi, ops

    steps = CONV_CHAINS[chain]
    src = rand_img(size[0], size[1], dt, 70 + chain, alpha="mixed")
    n0 = capi.launch_count()
    got = ops.pointwise_chain(to_dev(src), steps)
    launches = capi.launch_count() - n0
    seq = to_dev(src)
    for st in steps:
        if st[0] == "blur":
            seq = ops.blur(seq, st[1])
        elif st[0] == "unsharp":
            seq = ops.unsharp_mask(seq, st[1], st[2], st[3], st[4])
        else:
            seq = _single(ops, seq, st)
    assert_exact(got, to_host(seq), f"conv chain {chain} vs per-node kernels")
    aligned = (size[0] * 4 * np.dtype(dt).itemsize) % 16 == 0
    assert launches == (1 if (chain < 4 and aligned) else 2), launches


def test_pointwise_chain_rejects():
    from toucan_b200 import capi, ops

    t = to_dev(rand_img(16, 8, np.float32, 1))
    with pytest.raises(capi.TcError):
        ops.pointwise_chain(t, [("invert",)] * 9)
    with pytest.raises(capi.TcError):
        ops.pointwise_chain(t, [])


@pytest.mark.parametrize("dt", NP_TYPES)
@pytest.mark.parametrize("premult", [False, True])
def test_over_constant_background_equals_over_a_filled_frame(dt, premult):
    """tc_over_fill: the Fill generator under a track stack as a colour.  Bit-identical to rendering the u8 Fill frame and
    compositing over it (the colour takes the generator's u8 quantisation)."""
    from toucan_b200 import ops

    w, h = 203, 77
    fg = rand_img(w, h, dt, 60, alpha="mixed")
    for color in ([0.0, 0.0, 0.0, 1.0], [0.3, 0.7, 0.123, 0.5], [1.0, 1.0, 1.0, 0.0]):
        frame = ops.over(to_dev(fg), ops.fill(w, h, color), premult_fg=premult)
        assert_exact(ops.over_fill(to_dev(fg), color, premult_fg=premult), to_host(frame), f"over_fill {color}")
        assert_exact(frame, O.comp(fg, O.fill(w, h, color, O.U8), premult), "over vs oracle")


def test_frame_buffer_cache_reuses_across_streams_in_order():
    """tc_free keeps the buffer by exact size with an event on the freeing stream; tc_malloc on ANOTHER stream takes it
    over after waiting for that event, so work queued before the free still completes first."""
    import ctypes

    from toucan_b200 import capi

    lib = capi.lib()
    w, h = 2048, 1024
    nbytes = w * h * 4 * 4
    s1, s2 = ctypes.c_void_p(), ctypes.c_void_p()
    capi.check(lib.tc_stream_create(ctypes.byref(s1)), "stream")
    capi.check(lib.tc_stream_create(ctypes.byref(s2)), "stream")
    p1 = ctypes.c_void_p()
    capi.check(lib.tc_malloc(ctypes.byref(p1), nbytes, s1), "tc_malloc")
    img = capi.tc_image(p1.value, w, h, capi.F32)
    big = torch.rand((h, w, 4), device="cuda")
    torch.cuda.synchronize()
    # a long chain of work on stream 1 writing the buffer, then the free
    src = capi.tc_image(big.data_ptr(), w, h, capi.F32)
    for _ in range(20):
        capi.check(lib.tc_blur(ctypes.byref(img), ctypes.byref(src), ctypes.c_float(20.0), s1), "tc_blur")
    capi.check(lib.tc_free(p1, s1), "tc_free")
    p2 = ctypes.c_void_p()
    capi.check(lib.tc_malloc(ctypes.byref(p2), nbytes, s2), "tc_malloc")
    assert p2.value == p1.value  # same size: the cached buffer, not a new allocation
    img2 = capi.tc_image(p2.value, w, h, capi.F32)
    capi.check(lib.tc_fill(ctypes.byref(img2), capi.f4([0.25, 0.5, 0.75, 1.0]), s2), "tc_fill")
    out = torch.empty((h, w, 4), pin_memory=True)
    capi.check(lib.tc_download(ctypes.c_void_p(out.data_ptr()), p2, nbytes, s2), "tc_download")
    capi.check(lib.tc_stream_sync(s2), "sync")
    capi.check(lib.tc_stream_sync(s1), "sync")
    want = torch.tensor([0.25, 0.5, 0.75, 1.0]).expand(h, w, 4)
    assert torch.equal(out, want)  # the fill on stream 2 ran after stream 1's blurs, not under them
    capi.check(lib.tc_free(p2, s2), "tc_free")
    # another size does not match the cached buffer
    p3 = ctypes.c_void_p()
    capi.check(lib.tc_malloc(ctypes.byref(p3), nbytes // 2, s2), "tc_malloc")
    assert p3.value != p2.value
    capi.check(lib.tc_free(p3, s2), "tc_free")
    capi.check(lib.tc_pool_trim(), "tc_pool_trim")
    lib.tc_stream_destroy(s1)
    lib.tc_stream_destroy(s2)


# ---------------------------------------------------------------------------------------------------------
# compute-sanitizer is closed on this GPU pool (profiles/r02_compute_sanitizer_refused.txt), so the tiled kernels are
# checked the way the pool asks for: outputs inside guard bands that must stay untouched (an out-of-bounds store shows),
# sources followed by poisoned memory (an out-of-bounds load that reaches the result shows), and the same launch repeated
# many times with bit-identical results (a race between the asynchronous tile copies and the warps that read them would
# not reproduce).  Sizes are chosen to have partial tiles and border tiles on every side.
def _guarded(w, h, dt=torch.float32, fill=float("nan")):
    pad = 4096
    flat = torch.full((pad + h * w * 4 + pad,), fill, dtype=dt, device="cuda")
    return flat, flat[pad:pad + h * w * 4].view(h, w, 4)


@pytest.mark.parametrize("w,h,radius", [(203, 77, 10.0), (64, 24, 10.0), (129, 49, 20.0), (517, 263, 10.0)])
def test_stack_kernel_stays_inside_its_buffers_and_is_deterministic(T, w, h, radius):
    srcs = []
    for i in range(6):
        flat, view = _guarded(w, h)
        view.copy_(to_dev(rand_img(w, h, np.float32, 900 + i)))
        srcs.append((flat, view))
    d = [v for _, v in srcs]
    layers = [(d[0], None, radius, 0.0, 0.0), (d[1], d[2], radius, radius, 0.3), (d[3], None, 0.0, 0.0, 0.0), (d[4], d[5], radius, radius, 0.75)]
    flat, out = _guarded(w, h, fill=123.0)
    first = None
    for rep in range(25):
        out.fill_(-7.0)
        T.stack_f32(w, h, [0.1, 0.2, 0.3, 1.0], layers, out=out)
        torch.cuda.synchronize()
        assert bool((flat[:4096] == 123.0).all()) and bool((flat[-4096:] == 123.0).all()), "store outside the output frame"
        assert bool(torch.isfinite(out).all()), "a poisoned (out-of-frame) source value reached the result"
        if first is None:
            first = out.clone()
        else:
            assert torch.equal(first, out), f"run {rep} differs from run 0"
    a = [to_host(v) for v in d]
    acc = O.fill(w, h, [0.1, 0.2, 0.3, 1.0], O.F32)
    for sa, sb, ra, rb, v in [(a[0], None, radius, 0.0, 0.0), (a[1], a[2], radius, radius, 0.3), (a[3], None, 0.0, 0.0, 0.0), (a[4], a[5], radius, radius, 0.75)]:
        x = O.blur(sa, ra) if ra > 0 else sa
        if sb is not None:
            x = O.dissolve(x, O.blur(sb, rb) if rb > 0 else sb, v)
        acc = O.comp(x, acc, True)
    assert_close(first, acc, "guarded stack vs oracle")


@pytest.mark.parametrize("op", ["blur", "unsharp", "resize_down", "resize_up", "rotate"])
def test_tiled_filters_are_deterministic_with_poisoned_surroundings(T, op):
    w, h = 333, 217
    flat, src = _guarded(w, h)
    src.copy_(to_dev(rand_img(w, h, np.float32, 950)))
    first = None
    for rep in range(10):
        if op == "blur":
            got = T.blur(src, 10.0)
        elif op == "unsharp":
            got = T.unsharp_mask(src, "gaussian", 10.0, 1.0, 0.0)
        elif op == "resize_down":
            got = T.resize(src, (160, 101))
        elif op == "resize_up":
            got = T.resize(src, (700, 450))
        else:
            got = T.rotate(src, 33.0)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(got).all()), op
        if first is None:
            first = got.clone()
        else:
            assert torch.equal(first, got), f"{op}: run {rep} differs from run 0"


_SWITCH_PROBE = r"""
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, {root!r})
from toucan_b200 import capi, ops
capi.check(capi.lib().tc_set_device(0), "tc_set_device")
rng = np.random.default_rng(77)
src = torch.from_numpy(rng.random((217, 333, 4), dtype=np.float32)).cuda()
for name, out in (("rotate", ops.rotate(src, 33.0)), ("resize_down", ops.resize(src, (160, 101))), ("resize_up", ops.resize(src, (700, 450))),
                  ("blur", ops.blur(src, 10.0))):
    torch.cuda.synchronize()
    a = out.cpu().numpy()
    np.save(sys.argv[1] + "_" + name + ".npy", a)
    print(name, hashlib.sha256(a.tobytes()).hexdigest(), float(a.astype(np.float64).sum()))
"""


@pytest.mark.gpu
def test_kernel_variants_behind_the_switches_agree(T, tmp_path):
    """TOUCAN_ROTATE_TMA=0 (bounds-checked staging instead of the TMA copy), TOUCAN_RESIZE_WP=0 (the march instead of the
    warp-private kernel for float32 reductions) and TOUCAN_RESIZE_OCC=2 (its register-prefetch variant) change how pixels reach
    shared memory and which warp filters them, not the arithmetic: identical frames.  TOUCAN_CONV_TMA=0 selects round 1's Blur kernel:
    same taps, within float rounding; TOUCAN_ROTATE_QUAD=0 the one-pixel-per-thread Rotate: same weights, another order."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "probe.py"
    script.write_text(_SWITCH_PROBE.format(root=root))

    def run(env, tag):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, str(script), str(tmp_path / tag)], capture_output=True, text=True, timeout=300, env=e)
        assert r.returncode == 0, r.stderr[-2000:]
        return {l.split()[0]: (l.split()[1], float(l.split()[2])) for l in r.stdout.splitlines() if l and l.split()[0] in ("rotate", "resize_down", "resize_up", "blur")}

    base = run({"TOUCAN_ROTATE_QUAD": "0"}, "base")
    alt = run({"TOUCAN_ROTATE_QUAD": "0", "TOUCAN_ROTATE_TMA": "0", "TOUCAN_RESIZE_OCC": "2", "TOUCAN_RESIZE_WP": "0", "TOUCAN_CONV_TMA": "0"}, "alt")
    for k in ("rotate", "resize_down", "resize_up"):
        assert base[k][0] == alt[k][0], k
    assert abs(base["blur"][1] - alt["blur"][1]) <= 1e-6 * abs(base["blur"][1]), (base["blur"], alt["blur"])
    # the default Rotate of float32 frames (four pixels per thread, k_rotate_quad) sums the same weights in another order; its
    # per-pixel form for quads that miss the fast conditions is forced by TOUCAN_ROTATE_QUAD=generic
    run({}, "quad")
    run({"TOUCAN_ROTATE_QUAD": "generic"}, "generic")
    ref = np.load(str(tmp_path / "base_rotate.npy"))
    for tag in ("quad", "generic"):
        got = np.load(str(tmp_path / (tag + "_rotate.npy")))
        assert got.shape == ref.shape and float(np.abs(got - ref).max()) <= 2e-6, (tag, float(np.abs(got - ref).max()))


@pytest.mark.parametrize("seed", range(6))
def test_resize_random_geometries(T, seed):
    """Random source / destination sizes and storage types: every dispatch of tc_resize (the march with 1-4 staged spans, two
    and three CTAs per SM, chunk boundaries, the ring at 1-32 taps; the two-pass and gather paths beyond) against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        w, h = int(rng.integers(1, 420)), int(rng.integers(1, 420))
        dw, dh = int(rng.integers(1, 420)), int(rng.integers(1, 420))
        if rng.random() < 0.3:  # near-identity and exact integer ratios are where rounding of the tap positions bites
            dw, dh = max(1, w * int(rng.integers(1, 4)) // int(rng.integers(1, 4))), max(1, h * int(rng.integers(1, 4)) // int(rng.integers(1, 4)))
        dt = NP_TYPES[int(rng.integers(0, len(NP_TYPES)))]
        src = rand_img(w, h, dt, int(rng.integers(0, 1 << 30)))
        assert_close(T.resize(to_dev(src), (dw, dh)), O.resize(src, (dw, dh)), f"resize {np.dtype(dt).name} {(w, h)} -> {(dw, dh)}")


@pytest.mark.parametrize("seed", range(3))
def test_rotate_random_geometries(T, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(10):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        angle = float(rng.uniform(-360.0, 360.0)) if rng.random() < 0.8 else float(rng.choice([0.0, 90.0, 180.0, 270.0, 45.0, -45.0]))
        dt = NP_TYPES[int(rng.integers(0, len(NP_TYPES)))]
        src = rand_img(w, h, dt, int(rng.integers(0, 1 << 30)))
        assert_close(T.rotate(to_dev(src), angle), O.rotate(src, angle), f"rotate {np.dtype(dt).name} {(w, h)} by {angle}")


@pytest.mark.parametrize("w,h", [(64, 48), (332, 216), (4, 4), (36, 301), (258, 130)])
@pytest.mark.parametrize("dt", NP_TYPES)
def test_rotate_storage_type_tiles_through_the_copy_engine(T, dt, w, h):
    """Row pitches that are multiples of 16 bytes in every storage type: k_rotate_quad's tile arrives by TMA in the storage type
    (first column rounded down to a 16-byte boundary, converted once per staged pixel) -- partial blocks on every side, frames
    smaller than a block, several angles including the axis-aligned ones."""
    src = rand_img(w, h, dt, 4242 + w)
    for angle in (45.0, -17.3, 90.0, 180.0, 3.0, 211.0):
        assert_close(T.rotate(to_dev(src), angle), O.rotate(src, angle), f"rotate {np.dtype(dt).name} {(w, h)} by {angle}")


@pytest.mark.parametrize("seed", range(4))
def test_blur_and_unsharp_random_geometries(T, seed):
    """Random frame sizes (down to one pixel and narrower than a tile), radii over every tile template and storage types:
    the TMA kernel, round 1's tile kernel (odd row pitches, narrow types at large radii, frames smaller than a tile) and the
    wide-radius kernel against the oracle."""
    rng = np.random.default_rng(3000 + seed)
    for _ in range(8):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        radius = float(rng.choice([0.5, 1.0, 2.5, 4.0, 7.0, 10.0, 13.0, 17.5, 20.0, 26.0, 33.0]))
        dt = NP_TYPES[int(rng.integers(0, len(NP_TYPES)))]
        src = rand_img(w, h, dt, int(rng.integers(0, 1 << 30)))
        d = to_dev(src)
        what = f"{np.dtype(dt).name} {(w, h)} r={radius}"
        assert_close(T.blur(d, radius), O.blur(src, radius), "blur " + what)
        assert_close(T.unsharp_mask(d, "gaussian", radius, 0.8, 0.0), O.unsharp_mask(src, "gaussian", radius, 0.8, 0.0), "unsharp " + what)


@pytest.mark.parametrize("seed", range(3))
def test_crop_and_transitions_random_geometries(T, seed):
    """Crop boxes that leave the frame, transitions and composites between frames of different sizes and storage types
    (bit-exact ops: every pixel)."""
    rng = np.random.default_rng(4000 + seed)
    for _ in range(8):
        w, h = int(rng.integers(1, 260)), int(rng.integers(1, 200))
        dta, dtb = NP_TYPES[int(rng.integers(0, 4))], NP_TYPES[int(rng.integers(0, 4))]
        a, b = rand_img(w, h, dta, int(rng.integers(0, 1 << 30))), rand_img(w, h, dtb, int(rng.integers(0, 1 << 30)))
        da, db = to_dev(a), to_dev(b)
        v = float(rng.choice([0.0, 0.25, 1.0 / 3.0, 0.5, 0.9, 1.0]))
        what = f"{np.dtype(dta).name}/{np.dtype(dtb).name} {(w, h)} v={v}"
        assert_exact(T.dissolve(da, db, v), O.dissolve(a, b, v), "dissolve " + what)
        for vertical in (False, True):
            assert_exact(T.wipe(da, db, v, vertical), O.wipe(a, b, v, vertical), f"wipe {vertical} " + what)
        assert_exact(T.over(da, db, premult_fg=True), O.over(O.premult(a), b), "premult+over " + what)
        pos = (int(rng.integers(-20, w + 10)), int(rng.integers(-20, h + 10)))
        size = (int(rng.integers(1, w + 30)), int(rng.integers(1, h + 30)))
        assert_exact(T.crop(da, pos, size), O.crop(a, pos, size), f"crop {pos} {size} " + what)


@pytest.mark.parametrize("dt", NP_TYPES)
def test_chains_of_mirrorings_only(T, dt):
    """Flip + Flip (cancels: the copy each node makes), Flip + Flop (a copy read backwards along both axes), with and without a
    pointwise step between them -- found by the random timelines: the chain entry point refused chains without a pointwise step."""
    from toucan_b200 import ops

    src = rand_img(37, 23, dt, 123)
    d = to_dev(src)
    assert_exact(ops.pointwise_chain(d, [("flip",), ("flip",)]), src, "flip flip")
    assert_exact(ops.pointwise_chain(d, [("flop",), ("flop",)]), src, "flop flop")
    assert_exact(ops.pointwise_chain(d, [("flip",), ("flop",)]), O.flop(O.flip(src)), "flip flop")
    assert_exact(ops.pointwise_chain(d, [("flip",), ("invert",), ("flop",)]), O.flop(O.invert(O.flip(src))), "flip invert flop")


@pytest.mark.parametrize("seed", range(12))
def test_random_pointwise_chains_equal_the_nodes_one_by_one(seed):
    """Random chains (2 - 8 steps of every chainable kind, optionally behind a Blur / UnsharpMask head) over random sizes and
    storage types: the fused launch is bit-identical to the per-node kernels run in sequence."""
    from toucan_b200 import ops

    rng = np.random.default_rng(5000 + seed)
    for _ in range(6):
        dt = NP_TYPES[int(rng.integers(0, 4))]
        w, h = int(rng.integers(1, 200)), int(rng.integers(1, 120))
        pool = [("invert",), ("pow", float(rng.choice([2.0, 3.0, 2.2, 0.45]))), ("saturate", float(rng.uniform(0, 2))),
                ("color_map", str(rng.choice(["plasma", "viridis", "no such map"]))), ("premult",), ("unpremult",), ("flip",), ("flop",),
                ("over_fill", tuple(float(v) for v in rng.random(4)), bool(rng.integers(0, 2))),
                ("color_convert", "Linear Rec.709 (sRGB)", str(rng.choice(["ACEScct", "sRGB - Texture", "no such space"])), bool(rng.integers(0, 2)))]
        steps = [pool[int(rng.integers(0, len(pool)))] for _ in range(int(rng.integers(2, 8)))]
        src = rand_img(w, h, dt, int(rng.integers(0, 1 << 30)), alpha="mixed")
        got = ops.pointwise_chain(to_dev(src), steps)
        seq = to_dev(src)
        for st in steps:
            seq = _single(ops, seq, st)
        assert_exact(got, to_host(seq), f"{np.dtype(dt).name} {(w, h)} {steps}")
        # a convolution node at the head: the simple followers ride in its epilogue, anything else runs over a temporary
        head = ("blur", float(rng.choice([2.0, 10.0, 20.0]))) if rng.random() < 0.5 else ("unsharp", "gaussian", 3.0, 0.7, 0.0)
        got = ops.pointwise_chain(to_dev(src), [head] + steps)
        seq = ops.blur(to_dev(src), head[1]) if head[0] == "blur" else ops.unsharp_mask(to_dev(src), head[1], head[2], head[3], head[4])
        for st in steps:
            seq = _single(ops, seq, st)
        assert_exact(got, to_host(seq), f"{np.dtype(dt).name} {(w, h)} {[head] + steps}")


@pytest.mark.parametrize("seed", range(8))
def test_random_stacks_against_the_oracle(T, seed):
    """tc_stack_f32 on random stacks: 1 - 14 layers, each a plain, blurred or dissolving (equal or different blurs on its two
    inputs) float frame, random frame sizes from a few pixels (the fallback kernel) to several tiles (the TMA kernel)."""
    rng = np.random.default_rng(6000 + seed)
    w, h = int(rng.integers(3, 400)), int(rng.integers(3, 260))
    if seed % 3 == 0:
        w, h = (w + 63) // 64 * 64, (h + 23) // 24 * 24  # whole tiles
    n = int(rng.integers(1, 15))
    imgs = [rand_img(w, h, np.float32, int(rng.integers(0, 1 << 30)), alpha="mixed") for _ in range(min(2 * n, 12))]
    layers, spec = [], []
    for i in range(n):
        a = imgs[int(rng.integers(0, len(imgs)))]
        kind = int(rng.integers(0, 4))
        ra = float(rng.choice([0.0, 4.0, 10.0, 20.0]))
        if kind == 0:
            spec.append((a, None, ra, 0.0, 0.0))
        else:
            b = imgs[int(rng.integers(0, len(imgs)))]
            rb = ra  # (a Dissolve of differently blurred inputs is not a fusable layer: the host leaves it to the node path)
            spec.append((a, b, ra, rb, float(rng.choice([0.0, 0.25, 0.5, 1.0]))))
    bgcol = [float(v) for v in rng.random(4)]
    dev = {id(a): to_dev(a) for a in imgs}
    for a, b, ra, rb, v in spec:
        layers.append((dev[id(a)], dev[id(b)] if b is not None else None, ra, rb, v))
    got = T.stack_f32(w, h, bgcol, layers)
    acc = O.fill(w, h, bgcol, O.F32)
    for a, b, ra, rb, v in spec:
        x = O.blur(a, ra) if ra > 0 else a
        if b is not None:
            x = O.dissolve(x, O.blur(b, rb) if rb > 0 else b, v)
        acc = O.comp(x, acc, True)
    assert_close(got, acc, f"stack {(w, h)} {[(ra, rb, v, b is not None) for a, b, ra, rb, v in spec]}")


def test_color_convert_every_pair_of_spaces(T):
    """Every ordered pair of the colour spaces the baked path knows (curve -> 3x3 -> curve), float32, with and without the
    unpremult flag, against the oracle (relative bar where a decode curve reaches large values)."""
    from oracle.oracle import COLOR_SPACES

    rng = np.random.default_rng(31)
    src = rng.random((37, 53, 4), dtype=np.float32)
    src[..., 3] = np.where(rng.random((37, 53)) < 0.3, 1.0, src[..., 3])
    d = to_dev(src)
    names = sorted(COLOR_SPACES)
    for a in names:
        for b in names:
            for un in (False, True):
                # (the ACEScc / ACEScct decodes are 2^(17.52 x - 9.72): a float ulp of that exponent is 1e-6 of the result, and dividing by
                # a small alpha first sends x far above 1)
                rtol = 2e-5 if a.startswith("ACEScc") else 2e-6
                assert_close(T.colorconvert(d, a, b, unpremult=un), O.colorconvert(src, a, b, unpremult=un), f"{a} -> {b} unpremult={un}", tol=2e-5, rtol=rtol)


@pytest.mark.parametrize("seed", range(4))
def test_generators_random_parameters_bit_exact(T, seed):
    """Fill / Checkers / Gradient / Noise with random sizes, colours (also outside [0, 1]), checker sizes, noise kinds and seeds,
    in every storage type: bit-exact (north_star's bar for the generators and the seeded Noise)."""
    rng = np.random.default_rng(12000 + seed)
    types = [O.U8, O.U16, O.F16, O.F32]
    for _ in range(10):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        t = types[int(rng.integers(0, 4))]
        c1 = [float(v) for v in rng.uniform(-0.2, 1.3, 4)]
        c2 = [float(v) for v in rng.uniform(-0.2, 1.3, 4)]
        what = f"{(w, h)} type {t}"
        assert_exact(T.fill(w, h, c1, t), O.fill(w, h, c1, t), "fill " + what)
        cw, ch = int(rng.integers(1, 120)), int(rng.integers(1, 120))
        assert_exact(T.checkers(w, h, cw, ch, c1, c2, t), O.checkers(w, h, cw, ch, c1, c2, t), f"checkers {cw}x{ch} " + what)
        vertical = bool(rng.integers(0, 2))
        assert_exact(T.gradient(w, h, c1, c2, vertical, t), O.gradient(w, h, c1, c2, vertical, t), f"gradient {vertical} " + what)
        kind = str(rng.choice(["gaussian", "uniform", "white", "salt", "blue"]))
        a, b = float(rng.uniform(0, 1)), float(rng.uniform(0, 1))
        mono, sd = bool(rng.integers(0, 2)), int(rng.integers(0, 1000))
        assert_exact(T.noise(w, h, kind, a, b, mono, sd, t), O.noise(w, h, kind, a, b, mono, sd, t), f"noise {kind} {a} {b} {mono} {sd} " + what)


@pytest.mark.parametrize("dt", [np.float32, np.float16])
def test_pointwise_ops_on_special_values(T, dt):
    """Infinities, NaN, negative zero, denormals, negative and > 1 values through the pointwise ops in float storage: same bits as
    the oracle for the exact ops (NaN patterns compared as NaN), same NaN / infinity pattern and tolerance for the others."""
    rng = np.random.default_rng(77)
    special = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 6e-8, 65504.0, -65504.0, 0.5, 2.0, 1e-7, 3.0], dtype=np.float32)
    src = rng.choice(special, size=(24, 32, 4)).astype(dt)
    src2 = rng.choice(special, size=(24, 32, 4)).astype(dt)
    d, d2 = to_dev(src), to_dev(src2)

    def same(got, want, what, exact=True):
        got = to_host(got)
        assert got.dtype == want.dtype and got.shape == want.shape, what
        gn, wn = np.isnan(got.astype(np.float32)), np.isnan(want.astype(np.float32))
        assert np.array_equal(gn, wn), what + ": NaN pattern"
        if exact:
            assert np.array_equal(got[~gn].view(np.uint16 if dt == np.float16 else np.uint32), want[~wn].view(np.uint16 if dt == np.float16 else np.uint32)), what
        else:
            fin = np.isfinite(want.astype(np.float32)) & ~wn
            assert np.array_equal(np.isinf(got.astype(np.float32)), np.isinf(want.astype(np.float32))), what + ": infinities"
            assert np.allclose(got[fin].astype(np.float64), want[fin].astype(np.float64), rtol=2e-3 if dt == np.float16 else 1e-5, atol=1e-5), what

    with np.errstate(all="ignore"):
        same(T.invert(d), O.invert(src), "invert")
        same(T.premult(d), O.premult(src), "premult")
        same(T.flip(d), O.flip(src), "flip")
        same(T.dissolve(d, d2, 0.25), O.dissolve(src, src2, 0.25), "dissolve")
        same(T.wipe(d, d2, 0.5, False), O.wipe(src, src2, 0.5, False), "wipe")
        same(T.over(d, d2, premult_fg=True), O.over(O.premult(src), src2), "premult + over")
        same(T.saturate(d, 0.3), O.saturate(src, 0.3), "saturate", exact=False)
        same(T.unpremult(d), O.unpremult(src), "unpremult", exact=False)

"""In-tree build of the native parts (no cmake, no torch extension machinery).

  toucan_b200/lib/libtoucan_cuda.so   CUDA kernels + C ABI (include/toucan_cuda.h), sm_100a only
  toucan_b200/lib/libtoucan_render.so C++ host: ImageGraph / ImageEffectHost / OFX host (include/toucan_render.h)
  toucan_b200/lib/plugins/*.ofx       the six OpenFX plug-in modules
  toucan_b200/lib/toucan-render       command line renderer

nvcc cross-compiles for sm_100a without a GPU.  Artefacts are git-ignored but travel to
the GPU box with the repository snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "toucan_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB = os.path.join(PKG, "lib")
OBJ = os.path.join(LIB, "obj")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-ccbin", CXX, "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
                      "-I", os.path.join(ROOT, "include")]
# per-file flags: bit-exact files must not contract a*b+c
CU_FILES = {
    "tc_api.cu": [],
    "tc_pointwise.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "tc_resample.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
    "tc_conv.cu": [],
    "tc_stack.cu": [],
    "tc_y4m.cu": [],
    "tc_draw.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true"],
}
CPP_KERNEL_FILES = ["tc_tables.cpp"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unknown-pragmas", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
             "-I", HOST]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build command failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        sys.stderr.write(r.stdout + r.stderr)


def _headers(d):
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith((".h", ".cuh", ".hpp", ".inc"))]
    return out


def build_cuda(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers(CSRC) + _headers(os.path.join(ROOT, "include"))
    jobs = []
    objs = []
    for f, extra in CU_FILES.items():
        src = os.path.join(CSRC, f)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([NVCC] + NVCC_COMMON + extra + ["-c", src, "-o", obj])
    for f in CPP_KERNEL_FILES:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([CXX] + CXX_FLAGS + ["-c", src, "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    out = os.path.join(LIB, "libtoucan_cuda.so")
    if force or jobs or _newer(out, objs):
        _run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-ccbin", CXX, "-o", out] + objs, verbose)
    return out


HOST_LIB_SOURCES = ["otio.cpp", "PropertySet.cpp", "Plugin.cpp", "Util.cpp", "Image.cpp", "ImageNode.cpp", "ImageEffect.cpp",
                    "ImageEffectHost.cpp", "Comp.cpp", "StackFusion.cpp", "Read.cpp", "PngRead.cpp", "JpegRead.cpp", "ExrRead.cpp", "TiffRead.cpp", "Archive.cpp", "TimelineWrapper.cpp", "ImageGraph.cpp", "FrameScheduler.cpp", "render_capi.cpp"]
PLUGIN_MODULES = {
    "toucanGeneratorPlugin.ofx": ["GeneratorPlugin.cpp"],
    "toucanFilterPlugin.ofx": ["FilterPlugin.cpp"],
    "toucanTransformPlugin.ofx": ["TransformPlugin.cpp"],
    "toucanTransitionPlugin.ofx": ["TransitionPlugin.cpp"],
    "toucanColorPlugin.ofx": ["ColorPlugin.cpp"],
    "toucanDrawPlugin.ofx": ["DrawPlugin.cpp"],
}


def build_host(verbose=False, force=False):
    """C++ host library, OFX plug-in modules and the CLI.  Needs libtoucan_cuda.so."""
    src_dir = os.path.join(HOST, "toucanRender")
    if not os.path.isdir(src_dir) or not all(os.path.exists(os.path.join(src_dir, f)) for f in HOST_LIB_SOURCES):
        return None  # host tree incomplete: nothing to build yet
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.join(LIB, "plugins"), exist_ok=True)
    hdrs = _headers(HOST) + _headers(os.path.join(ROOT, "include"))
    jobs, objs = [], []
    for f in HOST_LIB_SOURCES:
        src = os.path.join(src_dir, f)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, "host_" + f + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([CXX] + CXX_FLAGS + ["-c", src, "-o", obj])
    plug_dir = os.path.join(HOST, "plugins")
    plug_common = [os.path.join(plug_dir, f) for f in ("Plugin.cpp", "Util.cpp")]
    plug_objs = {}
    for f in set(sum(PLUGIN_MODULES.values(), [])) | {"Plugin.cpp", "Util.cpp"}:
        src = os.path.join(plug_dir, f)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, "plug_" + f + ".o")
        plug_objs[f] = obj
        if force or _newer(obj, [src] + hdrs):
            jobs.append([CXX] + CXX_FLAGS + ["-I", plug_dir, "-c", src, "-o", obj])
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    rpath = ["-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/.."]
    out = os.path.join(LIB, "libtoucan_render.so")
    if objs and (force or jobs or _newer(out, objs)):
        _run([CXX, "-shared", "-o", out] + objs + ["-L", LIB, "-ltoucan_cuda", "-lz", "-ldl", "-lpthread"] + rpath, verbose)
    for mod, files in PLUGIN_MODULES.items():
        mobjs = [plug_objs[f] for f in files + ["Plugin.cpp", "Util.cpp"] if f in plug_objs]
        if len(mobjs) < 3:
            continue
        target = os.path.join(LIB, "plugins", mod)
        if force or jobs or _newer(target, mobjs):
            _run([CXX, "-shared", "-o", target] + mobjs + ["-L", LIB, "-ltoucan_cuda"] + rpath, verbose)
    main_src = os.path.join(HOST, "toucan-render", "main.cpp")
    if os.path.exists(main_src):
        exe = os.path.join(LIB, "toucan-render")
        if force or jobs or _newer(exe, [main_src, out] + hdrs):
            _run([CXX] + CXX_FLAGS + [main_src, "-o", exe, "-L", LIB, "-ltoucan_render", "-ltoucan_cuda", "-lz", "-ldl", "-lpthread"] + rpath, verbose)
    return out


def build_all(verbose=False, force=False):
    build_cuda(verbose, force)
    build_host(verbose, force)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built:", sorted(os.listdir(LIB)))

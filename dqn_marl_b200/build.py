"""In-tree build of libmarl_b200.so (nvcc, sm_100a only).

    python -m dqn_marl_b200.build [--force] [--verbose]

Every translation unit is compiled with
    -gencode arch=compute_100a,code=sm_100a -lineinfo
The env kernels additionally get -fmad=false: the reference's scores and rewards are sequences of
separately rounded float64 operations (people.py:287-291, evacuation_env.py:174-288) and a fused
multiply-add would change their bits.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
OUT = os.path.join(HERE, "libmarl_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
          "-I", os.path.join(HERE, "..", "include")]
# per-file extra flags
EXTRA = {
    "env.cu": ["-fmad=false", "-Xptxas", "-v"],
    "replay.cu": ["-Xptxas", "-v"],
    "qnet.cu": ["-Xptxas", "-v"],
    "gemm_tc.cu": ["-Xptxas", "-v"],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmarl_b200.so cannot be built (there is no CPU fallback)")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp", ".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "marl_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(OBJ, src + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        p = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, p

    objs = []
    logs = []
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:      # translation units are independent
        results = list(pool.map(compile_one, sources()))
    for src, obj, p in results:
        logs.append(f"== {src}\n{p.stdout}{p.stderr}")
        if p.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError(f"nvcc failed on {src}")
        if verbose:
            print(p.stdout + p.stderr)
        objs.append(obj)
    cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("nvcc link failed")
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(logs))
    return OUT


def build_env_trace() -> str:
    """libmarl_b200_envtrace.so: the same library with env.cu compiled -DMQ_ENV_TRACE (clock64 marks at the phase boundaries of
    env_step_kernel, read by scripts/env_phase_trace.py through MARL_B200_SO).  A profiling aid, never loaded by default."""
    build()
    nvcc = _nvcc()
    out = os.path.join(HERE, "libmarl_b200_envtrace.so")
    obj = os.path.join(OBJ, "env.cu.trace.o")
    cmd = [nvcc, *ARCH, *COMMON, "-fmad=false", "-DMQ_ENV_TRACE", "-c", os.path.join(CSRC, "env.cu"), "-o", obj]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    objs = [os.path.join(OBJ, src + ".o") for src in sources() if src != "env.cu"] + [obj]
    subprocess.run([nvcc, *ARCH, "-shared", "-o", out, *objs], check=True, capture_output=True, text=True)
    return out


def build_variant(name: str, defs) -> str:
    """libmarl_b200_<name>.so: the library with env.cu compiled with extra -D definitions (A/B timing of tuning knobs through
    MARL_B200_SO, e.g. build_variant("u1", ["-DMQ_SCORE_U8=1"]))."""
    build()
    nvcc = _nvcc()
    out = os.path.join(HERE, f"libmarl_b200_{name}.so")
    obj = os.path.join(OBJ, f"env.cu.{name}.o")
    subprocess.run([nvcc, *ARCH, *COMMON, "-fmad=false", *defs, "-c", os.path.join(CSRC, "env.cu"), "-o", obj], check=True, capture_output=True, text=True)
    objs = [os.path.join(OBJ, src + ".o") for src in sources() if src != "env.cu"] + [obj]
    subprocess.run([nvcc, *ARCH, "-shared", "-o", out, *objs], check=True, capture_output=True, text=True)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], sys.argv[k + 2:]))
    elif "--env-trace" in sys.argv:
        print(build_env_trace())
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

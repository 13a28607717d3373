"""Builds libb200zk.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os, subprocess, sys, hashlib, json

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200zk.so")
SOURCES = ["ntt.cu", "merkle.cu", "evaluator.cu", "msm.cu", "merkle_big.cu", "fr_ntt.cu", "lookup.cu", "stark.cpp", "verify.cpp", "capi.cpp", "timing.cpp", "jit.cpp", "poseidon_host.cpp", "groth16.cu", "c12_exec.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O3,-Wall",
         "-ccbin", "/usr/bin/g++", "-x", "cu"] + os.environ.get("B200_EXTRA_NVCC", "").split()


def _stamp():
    h = hashlib.sha256()
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            h.update(f.encode()); h.update(open(os.path.join(root, f), "rb").read())
    h.update(open(os.path.join(os.path.dirname(HERE), "include", "b200zk.h"), "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp_file = os.path.join(HERE, "build", "stamp.json")
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    st = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and json.load(open(stamp_file)).get("stamp") == st:
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "build", s + ".o")
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    ok = True
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (s, out))
        ok &= p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl", "-ccbin", "/usr/bin/g++"])
    json.dump({"stamp": st}, open(stamp_file, "w"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

"""CPU-side checks of the drop-in boundary: libmahout_b200.so builds, loads, exports every
symbol include/*.h declares, and refuses to run without a B200 (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mahout_b200 import build
    path = build.build()
    return C.CDLL(path)


def _declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.mb200_create(0, C.byref(h))
    assert rc == -6  # MB200_ERR_NO_DEVICE
    lib.mb200_last_error.restype = C.c_char_p
    assert b"no CPU fallback" in lib.mb200_last_error(None)


def test_product_does_not_touch_oracle():
    """The product path must never import, link or call oracle/."""
    pkg = os.path.join(ROOT, "mahout_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "mahout_oracle" not in text and "orc_" not in text, f


def test_hash_params_host_side(lib):
    """mb200_hash_params is host arithmetic (java.util.Random): same anchors as the oracle."""
    a = np.zeros(4, np.int64)
    b = np.zeros(4, np.int64)
    assert lib.mb200_hash_params(C.c_int64(42), 4, a.ctypes.data_as(C.c_void_p),
                                 b.ctypes.data_as(C.c_void_p)) == 0
    assert a.tolist() == [5025562857975149833, 5694868678511409995,
                          6169532649852302182, 6802844026563419272]
    assert b.tolist() == [5843495416241995736, 5111195811822994797,
                          1782466964123969572, 5086654115216342560]
    import oracle as orc
    for seed in (0, -7, 2 ** 62 + 12345):
        oa, ob = orc.hash_params(seed, 6)
        a = np.zeros(6, np.int64)
        b = np.zeros(6, np.int64)
        lib.mb200_hash_params(C.c_int64(seed), 6, a.ctypes.data_as(C.c_void_p),
                              b.ctypes.data_as(C.c_void_p))
        assert (a == oa).all() and (b == ob).all()


def test_cm_dims_errors(lib):
    w, d = C.c_int32(), C.c_int32()
    lib.mb200_cm_dims.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    assert lib.mb200_cm_dims(0.05, 0.01, C.byref(w), C.byref(d)) == 0
    import oracle as orc
    assert (w.value, d.value) == orc.cm_dims(0.05, 0.01)
    assert lib.mb200_cm_dims(0.0, 0.01, C.byref(w), C.byref(d)) == -7
    assert lib.mb200_cm_dims(0.5, 0.01, C.byref(w), C.byref(d)) == -7   # > 1/e
    assert lib.mb200_cm_dims(0.05, 3.0, C.byref(w), C.byref(d)) == -8   # > e
    assert lib.mb200_cm_dims(0.05, -1.0, C.byref(w), C.byref(d)) == -8


def test_hash_fold_matches_bigint_on_host():
    """cm_hash.cuh compiles for the host too: check the 2^63 == 25 (mod p) folding against Python
    big integers on adversarial inputs (no GPU needed)."""
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include <stdlib.h>
#include "cm_hash.cuh"
int main(int argc, char** argv) {
  long long a, b, k; unsigned w;
  while (scanf("%lld %lld %lld %u", &a, &b, &k, &w) == 4) {
    unsigned wm = (w > 1 && (w & (w - 1)) == 0) ? w - 1 : 0;
    printf("%u\n", cmh_column(cmh_residue(a), cmh_residue(b), cmh_residue(k), w, wm));
  }
  return 0;
}'''
    P = 2 ** 63 - 25
    rng = np.random.Generator(np.random.PCG64(7))
    cases = []
    edge = [0, 1, -1, 2 ** 63 - 1, -2 ** 63, P, P - 1, P + 1, -P, -P - 1, 25, -25, 2 ** 62]
    for a in edge:
        for k in edge:
            cases.append((a, (a * 7 + 3) % (2 ** 63), k, 1 << 20))
            cases.append((a, 2 ** 63 - 1, k, 4099))
    for _ in range(2000):
        a, b, k = (int(x) for x in rng.integers(-2 ** 63, 2 ** 63 - 1, 3, dtype=np.int64))
        cases.append((a, b, k, int(rng.integers(1, 2 ** 31 - 1))))
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "h.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "h")
        subprocess.check_call(["g++", "-O1", "-I", os.path.join(ROOT, "mahout_b200", "csrc"), cpp, "-o", exe])
        inp = "\n".join(f"{a} {b} {k} {w}" for a, b, k, w in cases)
        out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    want = [((a * k + b) % P) % w for a, b, k, w in cases]
    assert [int(x) for x in out] == want


def test_jni_glue_type_checks():
    """jni/mahout_b200_jni.c cannot be built without a JDK; it must at least type-check against
    include/mahout_b200.h (a stub jni.h supplies the JNI declarations it uses)."""
    import subprocess
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-Wno-unused-parameter",
                           "-I", os.path.join(ROOT, "tests", "jni_stub"), "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "jni", "mahout_b200_jni.c")])


def test_native_cli_builds_and_rejects_bad_arguments(tmp_path):
    """mahout_b200_itemsimilarity (C++ host driver over the C ABI): AbstractJob's -1 on bad arguments, and no
    CPU fallback when there is no device."""
    import subprocess
    import torch
    from mahout_b200 import build
    build.build()
    exe = build.CLI_BIN
    assert os.path.exists(exe)
    p = tmp_path / "in.csv"
    p.write_text("1,2,1\n")
    out = str(tmp_path / "o")
    assert subprocess.run([exe]).returncode == 255                                         # missing required options
    assert subprocess.run([exe, "-i", str(p), "-o", out, "-s", "SIMILARITY_TANIMOTO_COEFFICIENT"]).returncode == 255
    assert subprocess.run([exe, "-i", str(p), "-o", out, "-s", "SIMILARITY_COSINE", "-m", "0"]).returncode == 255
    assert subprocess.run([exe, "-i", str(p), "-o", out, "-s", "SIMILARITY_COSINE", "--bogus", "1"]).returncode == 255
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "-i", str(p), "-o", out, "-s", "SIMILARITY_COSINE"], capture_output=True, text=True)
        assert r.returncode == 255 and "no CPU fallback" in r.stderr
        assert not os.path.exists(out)


def test_java_double_to_string():
    from mahout_b200.similarity import java_double_to_string as j
    assert [j(v) for v in (0.4472135954999579, 1.0, 0.001, 0.0009765625, 1e7, 9999999.0, 1e-10, 100.0, -3.25, 0.0)] == \
        ["0.4472135954999579", "1.0", "0.001", "9.765625E-4", "1.0E7", "9999999.0", "1.0E-10", "100.0", "-3.25", "0.0"]


def test_hash_folding_arithmetic_on_host(tmp_path):
    """cm_hash.cuh compiled for the host: both the general 128-bit fold and the small-key path equal
    BigInteger arithmetic (HashFunction.java:31-34), including negative keys / parameters and the corners."""
    import random
    import subprocess
    exe = str(tmp_path / "hash_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "host", "hash_host_check.cpp"), "-o", exe])
    P = 2 ** 63 - 25
    rnd = random.Random(7)
    cases = []
    edge = [0, 1, -1, 2 ** 32 - 1, 2 ** 32, 2 ** 31, P - 1, P, P + 1, 2 ** 63 - 1, -2 ** 63, -2 ** 63 + 1, 1682, 10 ** 12]
    for a in edge[:10]:
        for k in edge:
            cases.append((a, rnd.randrange(P), k, rnd.choice([1, 7, 4096, 2 ** 20, 1000003, 2 ** 31 - 1])))
    for _ in range(20000):
        k = rnd.choice([rnd.randrange(2 ** 32), rnd.randrange(-2 ** 63, 2 ** 63), rnd.randrange(2 ** 24)])
        cases.append((rnd.randrange(2 ** 63), rnd.randrange(2 ** 63), k, rnd.choice([4096, 2 ** 20, 1000003, 97])))
    inp = "".join(f"{a} {b} {k} {w}\n" for a, b, k, w in cases)
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(cases)
    for (a, b, k, w), got in zip(cases, out):
        assert int(got) == ((a * k + b) % P) % w, (a, b, k, w)

"""Build-time guards read from the SASS of the in-tree library (cuobjdump; no GPU needed).

ptxas contracts `mul.rn.f32x2` followed by `add.rn.f32x2` into one FFMA2 (exactness rules): k_recon sums its
packed products as fma(product, one, acc) with `one` a kernel parameter.  If a toolchain ever sees through that, the
products' FMUL2 disappear and the sum rounds once where the reference rounds twice — the GPU parity tests would catch
it on a B200; this test catches it where the library is built."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.environ.get("HJK_SASS_OBJ", os.path.join(ROOT, "hijiki_b200", "lib", "obj", "context.o"))


def _sass(function_substring):
    if not shutil.which("cuobjdump") or not os.path.exists(OBJ):
        pytest.skip("cuobjdump or the built object is not available")
    out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    fn, body = None, {}
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            body[fn] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)\s*(.*?);", ln)
        if m and fn:
            body[fn].append((m.group(1), m.group(2)))
    hits = [f for f in body if function_substring in f]
    assert hits, f"no function matching {function_substring}"
    return body[hits[0]]


def test_packed_products_of_k_recon_are_not_contracted():
    ops = _sass("k_reconILb0ELi2ELb0")  # k_recon<false, 2, false>: the render path's instantiation
    last_writer = {}
    sums = 0
    for op, args in ops:
        regs = re.findall(r"\bR(\d+)\b", args)
        if op.startswith("FFMA2") and re.search(r"UR\d+\.F32\b", args):  # acc = product * one + acc
            product = int(regs[1])
            assert last_writer.get(product, "").startswith("FMUL2"), (
                f"the packed product in R{product} is written by {last_writer.get(product)!r}: contracted? ({op} {args})")
            sums += 1
        if regs and not op.startswith(("ST", "BRA", "BSSY", "BSYNC", "ISETP", "FSETP", "RED", "ATOM")):
            dst = int(regs[0])
            last_writer[dst] = op
            if "2" in op.split(".")[0][-1:] or ".64" in op or ".128" in op:  # register pairs / quads
                width = 4 if ".128" in op else 2
                for k in range(1, width):
                    last_writer[dst + k] = op
    assert sums >= 8, f"only {sums} packed accumulations found: has the kernel changed?"


@pytest.mark.parametrize("fn", ["k_trace_coopILi0E", "7k_traceILi0ELb0E"])
def test_trace_kernels_have_no_local_memory(fn):
    """The traversal stack lives in shared memory and the sphere-free trace kernels fit their registers: a spill here
    once cost a third of k_trace_coop through one unrelated statement (profiles/README.md, r02d)."""
    ops = _sass(fn)
    local = [op for op, _ in ops if op.startswith(("LDL", "STL"))]
    assert not local, f"{fn}: {len(local)} local-memory instructions"


def test_reconstruction_uses_the_tma_engine_and_packed_pairs():
    ops = [op for op, _ in _sass("k_reconILb0ELi2ELb0")]
    assert sum(op.startswith("UTMALDG") for op in ops) == 3  # layer 0, layer 1, the accumulator tile
    assert any(op.startswith("SYNCS") for op in ops)         # mbarrier
    assert sum(op.startswith("FFMA2") for op in ops) >= 20 and sum(op.startswith("FMUL2") for op in ops) >= 20
    assert not [op for op in ops if op.startswith(("LDL", "STL"))]

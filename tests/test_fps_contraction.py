"""The floating-point form of the FPS distance (oracle/fps.c header): nvcc's default -fmad=true contracts the
published sum of three squares to mul(y,y), fma(x,x,.), fma(z,z,.).  Checked here by compiling the expression and
tracing the registers through the PTX, so the claim the oracle and the CUDA kernels rest on stays reproducible."""
import re
import shutil
import subprocess

import pytest

SRC = r'''
extern "C" __global__ void dist(const float* p, const float* q, float* out) {
  float x1 = q[0], y1 = q[1], z1 = q[2];
  float x2 = p[0], y2 = p[1], z2 = p[2];
  float mag = x2 * x2 + y2 * y2 + z2 * z2;
  float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
  out[0] = mag; out[1] = d;
}
'''


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_nvcc_contracts_the_published_expression_to_mul_fma_fma(tmp_path):
    (tmp_path / "k.cu").write_text(SRC)
    subprocess.run(["nvcc", "-ptx", "-arch=sm_100a", "-o", str(tmp_path / "k.ptx"), str(tmp_path / "k.cu")], check=True)
    ins = re.findall(r"^\s*(ld\.global\.f32|sub\.f32|mul\.f32|fma\.rn\.f32|add\.f32|st\.global\.f32)\s+(.*);", (tmp_path / "k.ptx").read_text(), flags=re.M)
    assert not [op for op, _ in ins if op == "add.f32"]          # every addition was folded into an fma
    val = {}      # register -> symbolic value

    def regs(s):
        return [r.strip() for r in s.split(",")]
    loads = []
    for op, args in ins:
        a = regs(args)
        if op == "ld.global.f32":
            loads.append(a[0])
            val[a[0]] = ("q" if len(loads) <= 3 else "p") + "xyz"[(len(loads) - 1) % 3]   # q = picked point, p = candidate
        elif op == "sub.f32":
            val[a[0]] = f"({val[a[1]]}-{val[a[2]]})"
        elif op == "mul.f32":
            val[a[0]] = f"mul({val[a[1]]},{val[a[2]]})"
        elif op == "fma.rn.f32":
            val[a[0]] = f"fma({val[a[1]]},{val[a[2]]},{val[a[3]]})"
        elif op == "st.global.f32":
            val[a[0]] = val[a[1]]
    stored = [val[regs(args)[0]] for op, args in ins if op == "st.global.f32"]
    assert stored[0] == "fma(pz,pz,fma(px,px,mul(py,py)))", stored[0]
    assert stored[1] == "fma((pz-qz),(pz-qz),fma((px-qx),(px-qx),mul((py-qy),(py-qy))))", stored[1]

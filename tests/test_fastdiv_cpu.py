"""exact_div (draw_b200/csrc/device_math.cuh) == IEEE division: the three-operation quotient from a
correctly rounded reciprocal that k_tile uses under TRI_FASTDIV.  Checked on the CPU with the same
operations (fmaf is a single rounding in hardware with -mfma), including near-midpoint quotients."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_exact_div_matches_ieee_division(tmp_path):
    exe = tmp_path / "fastdiv_check"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(HERE, "native", "fastdiv_check.c"), "-lm"], check=True)
    out = subprocess.run([str(exe), "7", "3000000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    last = out.stdout.strip().splitlines()[-1]
    assert last.startswith("tested=") and last.endswith("bad=0"), last
    assert int(last.split()[0].split("=")[1]) > 30_000_000

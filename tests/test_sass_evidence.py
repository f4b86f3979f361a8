"""Static evidence that the built library's hot path is the Blackwell one (no GPU needed): cuobjdump -sass of
libmapnet_b200.so must contain the tcgen05 / TMA / TMEM instructions the conv engines are written around
(/opt/skills/guides/B200_PROFILING.md lists the mnemonics), for sm_100a, and none of the Hopper- or Ampere-style
tensor-core instructions a compatibility path would leave behind."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    from geomapnet_b200 import build
    so = build.build(verbose=False)
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    r = subprocess.run([exe, "-sass", so], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1000:]
    return r.stdout


def _kernel_bodies(sass):
    """{demangled-ish function name: SASS text}"""
    out, name, buf = {}, None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = "\n".join(buf)
            name, buf = m.group(1), []
        elif name:
            buf.append(line)
    if name:
        out[name] = "\n".join(buf)
    return out


def test_library_targets_sm_100a_only(sass):
    # every device function in the fat binary is sm_100a code (the CUDA runtime's own empty stub image is ignored)
    arch, per_arch = None, {}
    for line in sass.splitlines():
        m = re.match(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        elif re.match(r"\s*Function : ", line):
            per_arch[arch] = per_arch.get(arch, 0) + 1
    assert set(per_arch) == {"sm_100a"} and per_arch["sm_100a"] > 50, per_arch


def test_conv_engines_use_tcgen05_tma_and_tmem(sass):
    k = _kernel_bodies(sass)
    conv = {n: b for n, b in k.items() if "k_tc_conv" in n or "k_tc_wgrad" in n}
    assert len(conv) >= 20, sorted(conv)[:5]
    for n, b in conv.items():
        assert "UTCHMMA" in b, "no tcgen05.mma in %s" % n              # 5th-gen tensor core MMA, accumulator in TMEM
        assert "UTMALDG" in b, "no TMA tensor load in %s" % n          # cp.async.bulk.tensor
        assert "LDTM" in b, "no tcgen05.ld (TMEM read) in %s" % n
        assert "UTCBAR" in b, "no tcgen05.commit in %s" % n
        assert not re.search(r"\bHMMA\.|\bHGMMA|\bWGMMA", b), "legacy tensor-core instruction in %s" % n
    # the CTA-pair engines issue cta_group::2 MMAs and 2-CTA TMA loads
    pair = [b for n, b in conv.items() if "k_tc_conv2" in n or "k_tc_wgrad2" in n]
    assert pair and all("UTCHMMA.2CTA" in b and re.search(r"UTMALDG\.\dD\.2CTA", b) for b in pair)
    # im2col happens in the load: the fprop / dgrad / wgrad engines read activations through 4-D tensor maps
    assert all("UTMALDG.4D" in b for n, b in conv.items())
    # wgrad accumulates with fp32 reductions into the flat gradient buffer
    assert all(re.search(r"\bRED\b|REDG|ATOMG|RED\.", b) for n, b in conv.items() if "k_tc_wgrad" in n)


def test_bandwidth_kernels_are_vectorised(sass):
    k = _kernel_bodies(sass)
    for key in ("k_bn_apply", "k_bn_bwd_apply", "k_adam"):
        bodies = [b for n, b in k.items() if key in n]
        assert bodies, key
        for b in bodies:
            assert re.search(r"LDG\.E\.(128|64)|LDG\.E\.\w*\.?128", b), "%s has no 128-bit global loads" % key
            assert re.search(r"STG\.E\.(128|64)", b), "%s has no wide global stores" % key

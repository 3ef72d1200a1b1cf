"""The C-ABI library builds without a GPU, loads, and exports every symbol include/arp_b200.h declares.
No compute is invoked here (there is no GPU in the build container)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "arp_b200.h").read_text()
    return sorted(set(re.findall(r"ARP_API[^;(]*?\b(arp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    from arp_b200 import capi
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from {built_lib}"
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert exported == set(declared_symbols()), "library must export exactly the declared C ABI"
    header = (Path(__file__).resolve().parents[1] / "include" / "arp_b200.h").read_text()
    assert lib.arp_abi_version() == int(re.search(r"#define ARP_B200_ABI_VERSION (\d+)", header).group(1))


def test_sass_is_blackwell_native(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(built_lib)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):       # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, f"{mnemonic} not found in SASS"
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    from arp_b200 import capi
    with pytest.raises(capi.ArpError) as e:
        capi.Engine()
    assert e.value.code == capi.ARP_ERR_NO_DEVICE
    # and the drop-in entry point fails loudly too instead of computing on the CPU
    import numpy as np
    from arp_b200.label_reward import label_reward
    from arp_b200.store import NpyStore
    from arp_b200.synth import make_dataset, write_dataset
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        s = NpyStore(Path(d) / "ds", "w")
        write_dataset(s, make_dataset(n_episodes=2, len_lo=2, len_hi=4, size=64, num_frames=4))
        s.close()
        with pytest.raises(capi.ArpError):
            label_reward("coinrun", "hard", 500, 0, "x", d, data_path=str(Path(d) / "ds"),
                         clip_state_dict={"visual.proj": torch.zeros(768, 512)})

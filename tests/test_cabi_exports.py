"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for hdr in ("linne_encoder.h", "linne_decoder.h", "linne_b200.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(LINNE(?:Encoder|Decoder|B200)_\w+)\s*\(", text))
    return names


def test_headers_declare_the_reference_api():
    names = declared_functions()
    for n in ("LINNEEncoder_EncodeHeader", "LINNEEncoder_CalculateWorkSize", "LINNEEncoder_Create",
              "LINNEEncoder_Destroy", "LINNEEncoder_SetEncodeParameter", "LINNEEncoder_EncodeBlock",
              "LINNEEncoder_EncodeWhole", "LINNEDecoder_DecodeHeader", "LINNEDecoder_CalculateWorkSize",
              "LINNEDecoder_Create", "LINNEDecoder_Destroy", "LINNEDecoder_SetHeader",
              "LINNEDecoder_DecodeBlock", "LINNEDecoder_DecodeWhole"):
        assert n in names


def test_product_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "linne_b200", "liblinne_b200.so"))
    for name in sorted(declared_functions()):
        assert hasattr(lib, name), name
    lib.LINNEB200_Backend.restype = ctypes.c_char_p
    assert lib.LINNEB200_Backend() == b"cuda-sm_100a"


def _nm(path, *flags):
    import subprocess
    out = subprocess.run(["nm", "-D", *flags, path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_product_has_no_cpu_path_compiled_in():
    """The loop executor (tests/hostsim) must not be in the product: judged by the dynamic symbol tables, not by a string.
    The product defines the shim entry points and resolves them through the CUDA runtime (undefined cuda* symbols or a
    statically linked runtime with its device-code registration); the host simulator defines the same entry points and
    needs nothing from CUDA."""
    so = os.path.join(ROOT, "linne_b200", "liblinne_b200.so")
    sim = os.path.join(ROOT, "tests", "hostsim", "liblinne_hostsim.so")
    defined = _nm(so, "--defined-only")
    assert "lnb_shim_decode" in defined and "lnb_shim_hop" in defined
    assert not any("LoopExec" in s or "hostsim" in s.lower() for s in defined)
    blob = open(so, "rb").read()
    assert b".nv_fatbin" in blob or b"__nv_relfatbin" in blob or b"nv_fatbin" in blob      # device code is embedded
    assert b"hostsim" not in blob
    if os.path.exists(sim):
        sim_blob = open(sim, "rb").read()
        assert b"nv_fatbin" not in sim_blob                                           # ... and only there
        assert any("LoopExec" in s for s in _nm(sim, "--defined-only", "-C")) or b"LoopExec" in sim_blob


def test_every_declared_extension_symbol_is_exported():
    import re
    so = os.path.join(ROOT, "linne_b200", "liblinne_b200.so")
    defined = _nm(so, "--defined-only")
    for header in ("linne_b200.h", "linne_encoder.h", "linne_decoder.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        for name in set(re.findall(r"\b(LINNE(?:B200|Encoder|Decoder)_\w+)\s*\(", text)):
            assert name in defined, (header, name)


def test_struct_layouts_match_the_reference_abi():
    import harness
    # sizes/offsets printed by a C program compiled against the reference's include/ (x86-64 SysV)
    assert ctypes.sizeof(harness.LINNEHeader) == 36
    assert harness.LINNEHeader.num_samples.offset == 12 and harness.LINNEHeader.ch_process_method.offset == 32
    assert ctypes.sizeof(harness.LINNEEncodeParameter) == 20
    assert ctypes.sizeof(harness.LINNEEncoderConfig) == 16
    assert ctypes.sizeof(harness.LINNEDecoderConfig) == 16

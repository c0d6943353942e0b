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


def test_product_has_no_cpu_path_compiled_in():
    # the loop executor lives only in tests/hostsim; the product must not carry it
    blob = open(os.path.join(ROOT, "linne_b200", "liblinne_b200.so"), "rb").read()
    assert b"hostsim" not in blob


def test_struct_layouts_match_the_reference_abi():
    import harness
    # sizes/offsets printed by a C program compiled against the reference's include/ (x86-64 SysV)
    assert ctypes.sizeof(harness.LINNEHeader) == 36
    assert harness.LINNEHeader.num_samples.offset == 12 and harness.LINNEHeader.ch_process_method.offset == 32
    assert ctypes.sizeof(harness.LINNEEncodeParameter) == 20
    assert ctypes.sizeof(harness.LINNEEncoderConfig) == 16
    assert ctypes.sizeof(harness.LINNEDecoderConfig) == 16

"""Black-box replay of the reference's API-contract tests against our host code.

Sources: test/linne_encoder/linne_encoder_test.cpp (header encode + argument errors :49-122,
Create/CalculateWorkSize contracts :126-277, EncodeBlock errors :280-458) and
test/linne_decoder/linne_decoder_test.cpp (header round-trip + invalid headers :74-199, Create
contracts :203-349, corruption matrix :522-569).

Runs twice: on the CPU simulator of the product's host code + kernel bodies (`hostsim`, no GPU
needed) and, marked gpu, on the CUDA product through the C-ABI.
"""
import ctypes as C

import numpy as np
import pytest

import harness
from harness import (LINNEHeader, LINNEEncodeParameter, LINNEEncoderConfig, LINNEDecoderConfig,
                     OK, INVALID_ARGUMENT, INVALID_FORMAT, INSUFFICIENT_BUFFER, INSUFFICIENT_DATA,
                     PARAMETER_NOT_SET, DATA_CORRUPTION)


@pytest.fixture(params=["hostsim", "ref", pytest.param("gpu", marks=pytest.mark.gpu)])
def codec(request):
    """`ref` runs the same expectations on the unmodified reference (where oracle/_ref exists): it
    shows these tests state the reference's behaviour, not ours."""
    c = request.getfixturevalue(request.param)
    c.is_reference = request.param == "ref"
    return c


def skip_on_reference(codec, why):
    if codec.is_reference:
        pytest.skip("reference " + why)


def u8(buf):
    return buf.ctypes.data_as(C.POINTER(C.c_uint8))


def valid_header():
    return LINNEHeader(1, 2, 2, 44100, 44100, 16, 4096, 0, 1)


def test_encode_header_layout_and_errors(codec):
    L = codec.lib
    buf = np.zeros(64, np.uint8)
    assert L.LINNEEncoder_EncodeHeader(C.byref(valid_header()), u8(buf), 30) == OK
    want = b"IBRA" + (1).to_bytes(4, "big") + (2).to_bytes(4, "big") + (2).to_bytes(2, "big") + \
        (44100).to_bytes(4, "big") + (44100).to_bytes(4, "big") + (16).to_bytes(2, "big") + \
        (4096).to_bytes(4, "big") + bytes([0, 1])
    assert buf[:30].tobytes() == want
    assert L.LINNEEncoder_EncodeHeader(None, u8(buf), 30) == INVALID_ARGUMENT
    assert L.LINNEEncoder_EncodeHeader(C.byref(valid_header()), None, 30) == INVALID_ARGUMENT
    assert L.LINNEEncoder_EncodeHeader(C.byref(valid_header()), u8(buf), 29) == INSUFFICIENT_BUFFER
    for field, bad in (("num_channels", 0), ("num_samples", 0), ("sampling_rate", 0), ("bits_per_sample", 0),
                       ("num_samples_per_block", 0), ("preset", 8), ("ch_process_method", 2)):
        h = valid_header(); setattr(h, field, bad)
        assert L.LINNEEncoder_EncodeHeader(C.byref(h), u8(buf), 30) == INVALID_FORMAT, field
    h = valid_header(); h.num_channels = 1                      # M/S needs two channels
    assert L.LINNEEncoder_EncodeHeader(C.byref(h), u8(buf), 30) == INVALID_FORMAT


def test_decode_header_roundtrip_and_errors(codec):
    L = codec.lib
    buf = np.zeros(64, np.uint8)
    assert L.LINNEEncoder_EncodeHeader(C.byref(valid_header()), u8(buf), 30) == OK
    out = LINNEHeader()
    assert L.LINNEDecoder_DecodeHeader(u8(buf), 30, C.byref(out)) == OK
    for f, _ in LINNEHeader._fields_:
        assert getattr(out, f) == getattr(valid_header(), f), f
    assert L.LINNEDecoder_DecodeHeader(None, 30, C.byref(out)) == INVALID_ARGUMENT
    assert L.LINNEDecoder_DecodeHeader(u8(buf), 30, None) == INVALID_ARGUMENT
    assert L.LINNEDecoder_DecodeHeader(u8(buf), 29, C.byref(out)) == INSUFFICIENT_DATA
    bad = buf.copy(); bad[0] = ord("X")
    assert L.LINNEDecoder_DecodeHeader(u8(bad), 30, C.byref(out)) == INVALID_FORMAT


def test_set_header_validation(codec):
    L = codec.lib
    dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(2, 3, 128, 1)), None, 0)
    assert dec
    try:
        assert L.LINNEDecoder_SetHeader(dec, C.byref(valid_header())) == OK
        assert L.LINNEDecoder_SetHeader(None, C.byref(valid_header())) == INVALID_ARGUMENT
        assert L.LINNEDecoder_SetHeader(dec, None) == INVALID_ARGUMENT
        for field, bad in (("format_version", 0), ("codec_version", 3), ("num_channels", 0), ("num_samples", 0),
                           ("sampling_rate", 0), ("bits_per_sample", 0), ("num_samples_per_block", 0),
                           ("preset", 8), ("ch_process_method", 2)):
            h = valid_header(); setattr(h, field, bad)
            assert L.LINNEDecoder_SetHeader(dec, C.byref(h)) == INVALID_FORMAT, field
        h = valid_header(); h.num_channels = 3
        assert L.LINNEDecoder_SetHeader(dec, C.byref(h)) == INSUFFICIENT_BUFFER
    finally:
        L.LINNEDecoder_Destroy(dec)
    small = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(2, 2, 32, 1)), None, 0)
    try:
        h = valid_header(); h.preset = 7                        # needs 3 layers / 128 parameters
        assert L.LINNEDecoder_SetHeader(small, C.byref(h)) == INSUFFICIENT_BUFFER
    finally:
        L.LINNEDecoder_Destroy(small)


def test_work_size_and_create_contracts(codec):
    L = codec.lib
    ecfg = LINNEEncoderConfig(2, 4096, 3, 128)
    size = L.LINNEEncoder_CalculateWorkSize(C.byref(ecfg))
    assert size > 0
    assert L.LINNEEncoder_CalculateWorkSize(None) == -1
    for field in ("max_num_channels", "max_num_samples_per_block", "max_num_layers", "max_num_parameters_per_layer"):
        bad = LINNEEncoderConfig(2, 4096, 3, 128); setattr(bad, field, 0)
        assert L.LINNEEncoder_CalculateWorkSize(C.byref(bad)) == -1
        assert not L.LINNEEncoder_Create(C.byref(bad), None, 0)
    assert L.LINNEEncoder_CalculateWorkSize(C.byref(LINNEEncoderConfig(2, 64, 3, 128))) == -1   # params > block
    work = C.create_string_buffer(size)
    enc = L.LINNEEncoder_Create(C.byref(ecfg), work, size)
    assert enc
    L.LINNEEncoder_Destroy(enc)
    assert not L.LINNEEncoder_Create(C.byref(ecfg), work, size - 1)
    assert not L.LINNEEncoder_Create(C.byref(ecfg), None, size)
    assert not L.LINNEEncoder_Create(None, work, size)

    dcfg = LINNEDecoderConfig(2, 3, 128, 1)
    dsize = L.LINNEDecoder_CalculateWorkSize(C.byref(dcfg))
    assert dsize > 0 and L.LINNEDecoder_CalculateWorkSize(None) == -1
    for field in ("max_num_channels", "max_num_layers", "max_num_parameters_per_layer"):
        bad = LINNEDecoderConfig(2, 3, 128, 1); setattr(bad, field, 0)
        assert L.LINNEDecoder_CalculateWorkSize(C.byref(bad)) == -1
        assert not L.LINNEDecoder_Create(C.byref(bad), None, 0)
    dwork = C.create_string_buffer(dsize)
    dec = L.LINNEDecoder_Create(C.byref(dcfg), dwork, dsize)
    assert dec
    L.LINNEDecoder_Destroy(dec)
    assert not L.LINNEDecoder_Create(C.byref(dcfg), dwork, dsize - 1)


def test_set_encode_parameter_validation(codec):
    L = codec.lib
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(2, 4096, 3, 128)), None, 0)
    try:
        good = LINNEEncodeParameter(2, 16, 44100, 4096, 7, 1, 0, 0)
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(good)) == OK
        assert L.LINNEEncoder_SetEncodeParameter(None, C.byref(good)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_SetEncodeParameter(enc, None) == INVALID_ARGUMENT
        for field, bad in (("num_channels", 0), ("bits_per_sample", 0), ("sampling_rate", 0),
                           ("num_samples_per_block", 0), ("preset", 8), ("ch_process_method", 2),
                           ("num_samples_per_block", 128)):          # block must exceed every layer size
            p = LINNEEncodeParameter(2, 16, 44100, 4096, 7, 1, 0, 0); setattr(p, field, bad)
            assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(p)) == INVALID_FORMAT, field
        p = LINNEEncodeParameter(3, 16, 44100, 4096, 7, 0, 0, 0)
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(p)) == INSUFFICIENT_BUFFER
        p = LINNEEncodeParameter(2, 16, 44100, 8192, 7, 1, 0, 0)
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(p)) == INSUFFICIENT_BUFFER
    finally:
        L.LINNEEncoder_Destroy(enc)


def test_encode_block_argument_errors_and_silent_block(codec):
    L = codec.lib
    enc = L.LINNEEncoder_Create(C.byref(LINNEEncoderConfig(2, 1024, 3, 128)), None, 0)
    try:
        pcm = np.zeros((2, 1024), np.int32)
        ptrs = harness._chan_ptrs(pcm)
        out = np.zeros(1 << 16, np.uint8)
        size = C.c_uint32(0)
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1024, u8(out), len(out), C.byref(size)) == PARAMETER_NOT_SET
        assert L.LINNEEncoder_EncodeWhole(enc, ptrs, 1024, u8(out), len(out), C.byref(size)) == PARAMETER_NOT_SET
        prm = LINNEEncodeParameter(2, 16, 44100, 1024, 0, 1, 0, 0)
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(prm)) == OK
        assert L.LINNEEncoder_EncodeBlock(None, ptrs, 1024, u8(out), len(out), C.byref(size)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, None, 1024, u8(out), len(out), C.byref(size)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 0, u8(out), len(out), C.byref(size)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1024, None, len(out), C.byref(size)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1024, u8(out), 0, C.byref(size)) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1024, u8(out), len(out), None) == INVALID_ARGUMENT
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1025, u8(out), len(out), C.byref(size)) == INSUFFICIENT_BUFFER
        # digital silence -> an 11-byte SILENT block (linne_encoder_test.cpp silent-block check)
        assert L.LINNEEncoder_EncodeBlock(enc, ptrs, 1024, u8(out), len(out), C.byref(size)) == OK
        assert size.value == 11
        blk = out[:11].tobytes()
        assert blk[:2] == b"\xff\xff" and int.from_bytes(blk[2:6], "big") == 5 and blk[8] == 1
        assert int.from_bytes(blk[9:11], "big") == 1024
    finally:
        L.LINNEEncoder_Destroy(enc)


def test_output_buffer_too_small(codec):
    skip_on_reference(codec, "overruns the output buffer in Release instead (linne_encoder.c:699-749 has asserts only)")
    pcm = harness.synth_pcm(n=4096, channels=2, bits=16, seed=2)
    rc, _ = codec.encode(pcm, preset=0, block=2048, cap=30 + 600, return_code=True)
    assert rc == INSUFFICIENT_BUFFER
    rc, _ = codec.encode(pcm, preset=0, block=2048, cap=20, return_code=True)
    assert rc == INSUFFICIENT_BUFFER


def test_silent_stream_decode_size(codec, oracle):
    # test/linne_decoder/linne_decoder_test.cpp:395-467: silent block round trip
    pcm = np.zeros((2, 3000), np.int32)
    stream = codec.encode(pcm, preset=0, block=1024)
    assert len(stream) == 30 + 3 * 11
    assert stream == oracle.encode(pcm, preset=0, block=1024)
    assert np.array_equal(codec.decode(stream), pcm)


def test_decoder_corruption_matrix(codec, oracle):
    # test/linne_decoder/linne_decoder_test.cpp:522-569
    pcm = harness.synth_pcm(n=3000, channels=2, bits=16, seed=4)
    stream = oracle.encode(pcm, preset=0, block=1024)
    first_size = int.from_bytes(stream[32:36], "big") + 6

    def rc_of(data, crc=1):
        return codec.decode(bytes(data), check_crc=crc, return_code=True)[0]

    assert rc_of(stream) == OK
    for pos in (30, 31):                                    # sync code bytes
        bad = bytearray(stream); bad[pos] ^= 0xFF
        assert rc_of(bad) == INVALID_FORMAT
    for pos in (38, 39, 40, 30 + first_size - 1):           # type, sample count, last payload byte
        bad = bytearray(stream); bad[pos] ^= 0x01
        assert rc_of(bad) == DATA_CORRUPTION, pos
    assert rc_of(stream[:len(stream) - 1]) == INSUFFICIENT_DATA
    assert rc_of(stream[:30 + first_size + 5]) == INSUFFICIENT_DATA
    skip_on_reference(codec, "has no bounds checks on corrupt payloads when the CRC check is off")
    # with the CRC check off, a bad block type is a format error; a corrupt payload must not crash
    bad = bytearray(stream); bad[38] = 9
    assert rc_of(bad, crc=0) == INVALID_FORMAT
    bad = bytearray(stream)
    for i in range(60, 200): bad[i] ^= 0x5A
    assert rc_of(bad, crc=0) in (OK, DATA_CORRUPTION)
    # samples decoded before the failing block are delivered, the rest of the buffer is untouched
    bad = bytearray(stream); bad[30 + first_size + 20] ^= 0x01          # inside block 1
    rc, out = codec.decode(bytes(bad), return_code=True, fill=-7)
    assert rc == DATA_CORRUPTION
    assert np.array_equal(out[:, :1024], pcm[:, :1024]) and np.all(out[:, 1024:] == -7)


def test_decode_whole_buffer_checks(codec, oracle):
    pcm = harness.synth_pcm(n=2048, channels=2, bits=16, seed=6)
    stream = oracle.encode(pcm, preset=0, block=1024)
    assert codec.decode(stream, return_code=True, out_samples=2047)[0] == INSUFFICIENT_BUFFER
    assert codec.decode(stream, return_code=True, out_channels=1)[0] == INSUFFICIENT_BUFFER
    # a stream that ends at a block boundary before num_samples is OK (linne_decoder.c:708)
    first_size = int.from_bytes(stream[32:36], "big") + 6
    rc, out = codec.decode(stream[:30 + first_size], return_code=True)
    assert rc == OK and np.array_equal(out[:, :1024], pcm[:, :1024])


def test_decode_block_streaming(codec, oracle):
    # the player's loop: SetHeader once, then DecodeBlock per block (tools/linne_player/linne_player.c:110-121)
    L = codec.lib
    pcm = harness.synth_pcm(n=2500, channels=2, bits=16, seed=8)
    stream = oracle.encode(pcm, preset=2, block=1024)
    buf = np.zeros(len(stream) + 16, np.uint8); buf[:len(stream)] = np.frombuffer(stream, np.uint8)
    rc, hdr = codec.decode_header(stream)
    dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(2, 3, 128, 1)), None, 0)
    try:
        out = np.zeros((2, 1024), np.int32)
        used, got = C.c_uint32(0), C.c_uint32(0)
        assert L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + 30, C.POINTER(C.c_uint8)), len(stream) - 30,
                                          harness._chan_ptrs(out), 2, 1024, C.byref(used), C.byref(got)) == PARAMETER_NOT_SET
        assert L.LINNEDecoder_SetHeader(dec, C.byref(hdr)) == OK
        off, done = 30, 0
        while done < 2500:
            rc = L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + off, C.POINTER(C.c_uint8)), len(stream) - off,
                                            harness._chan_ptrs(out), 2, 1024, C.byref(used), C.byref(got))
            assert rc == OK
            assert used.value == int.from_bytes(stream[off + 2:off + 6], "big") + 6
            assert np.array_equal(out[:, :got.value], pcm[:, done:done + got.value])
            off += used.value; done += got.value
        assert off == len(stream)
        assert L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + 30, C.POINTER(C.c_uint8)), len(stream) - 30,
                                          harness._chan_ptrs(out), 2, 100, C.byref(used), C.byref(got)) == INSUFFICIENT_BUFFER
    finally:
        L.LINNEDecoder_Destroy(dec)


def _block_loop(L, dec, buf, size, channels, cap, start=30):
    """DecodeBlock until the stream ends or a call fails: [(rc, used, samples)], PCM [C][total]."""
    out = np.full((channels, cap), -7, np.int32)
    used, got = C.c_uint32(0), C.c_uint32(0)
    off, calls, pcm = start, [], []
    while off < size:
        rc = L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + off, C.POINTER(C.c_uint8)), size - off,
                                        harness._chan_ptrs(out), channels, cap, C.byref(used), C.byref(got))
        calls.append((rc, used.value, got.value))
        if rc != OK:
            break
        pcm.append(out[:, :got.value].copy())
        off += used.value
    return calls, (np.concatenate(pcm, axis=1) if pcm else np.zeros((channels, 0), np.int32))


@pytest.mark.parametrize("readahead", [2, 3, 16])
def test_decode_block_readahead_equals_batch_of_one(codec, oracle, readahead):
    """DecodeBlock with K blocks decoded per batch returns what the batch of one returns, call by
    call -- clean streams, mixed block types, a corrupted block in the middle, bytes changed between calls."""
    skip_on_reference(codec, "has no read-ahead")
    L = codec.lib
    cases = [(harness.synth_pcm(n=1024 * 7 + 300, channels=2, bits=16, seed=21), 16, 2, 1024),
             (harness.mixed_types_pcm(), 16, 5, 10240)]         # raw, silent and compressed blocks
    for pcm, bits, preset, block in cases:
        stream = oracle.encode(pcm, bits=bits, preset=preset, block=block)
        rc, hdr = codec.decode_header(stream)
        sizes, off = [], 30
        while off < len(stream):
            sizes.append(int.from_bytes(stream[off + 2:off + 6], "big") + 6); off += sizes[-1]
        variants = [bytearray(stream)]
        bad = bytearray(stream); bad[30 + sum(sizes[:3]) + 25] ^= 0x10; variants.append(bad)        # block 3 corrupted
        bad = bytearray(stream); bad[30 + sum(sizes[:2])] = 0x00; variants.append(bad)              # block 2 loses its sync code
        for image in variants:
            buf = np.zeros(len(image) + 16, np.uint8); buf[:len(image)] = np.frombuffer(bytes(image), np.uint8)
            results = []
            for k in (0, readahead):
                dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(pcm.shape[0], 3, 128, 1)), None, 0)
                try:
                    assert L.LINNEDecoder_SetHeader(dec, C.byref(hdr)) == OK
                    L.LINNEB200_DecoderSetReadahead(dec, k)
                    results.append(_block_loop(L, dec, buf, len(image), pcm.shape[0], block))
                finally:
                    L.LINNEDecoder_Destroy(dec)
            assert results[0][0] == results[1][0]
            assert np.array_equal(results[0][1], results[1][1])
        # the caller's bytes change after the batch was decoded: the cached copy must not be served
        buf = np.zeros(len(stream) + 16, np.uint8); buf[:len(stream)] = np.frombuffer(stream, np.uint8)
        dec = L.LINNEDecoder_Create(C.byref(LINNEDecoderConfig(pcm.shape[0], 3, 128, 1)), None, 0)
        try:
            assert L.LINNEDecoder_SetHeader(dec, C.byref(hdr)) == OK
            L.LINNEB200_DecoderSetReadahead(dec, readahead)
            out = np.zeros((pcm.shape[0], block), np.int32)
            used, got = C.c_uint32(0), C.c_uint32(0)
            assert L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + 30, C.POINTER(C.c_uint8)), len(stream) - 30,
                                              harness._chan_ptrs(out), pcm.shape[0], block, C.byref(used), C.byref(got)) == OK
            assert used.value == sizes[0] and np.array_equal(out[:, :got.value], pcm[:, :got.value])
            buf[30 + sizes[0] + 30] ^= 0x04
            assert L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + 30 + sizes[0], C.POINTER(C.c_uint8)),
                                              len(stream) - 30 - sizes[0], harness._chan_ptrs(out), pcm.shape[0], block,
                                              C.byref(used), C.byref(got)) == DATA_CORRUPTION
            # too small a buffer for a cached block: the ordinary path answers
            buf[30 + sizes[0] + 30] ^= 0x04
            assert L.LINNEDecoder_DecodeBlock(dec, C.cast(buf.ctypes.data + 30 + sizes[0], C.POINTER(C.c_uint8)),
                                              len(stream) - 30 - sizes[0], harness._chan_ptrs(out), pcm.shape[0], block - 1,
                                              C.byref(used), C.byref(got)) == INSUFFICIENT_BUFFER
        finally:
            L.LINNEDecoder_Destroy(dec)


def test_encode_block_loop_equals_encode_whole(codec):
    # the CLI encodes with EncodeHeader + EncodeBlock per block (tools/linne_codec/linne_codec.c:123-161)
    pcm = harness.synth_pcm(n=2500, channels=2, bits=16, seed=10)
    assert codec.encode(pcm, preset=2, block=1024, whole=False) == codec.encode(pcm, preset=2, block=1024, whole=True)


def test_block_longer_than_header_block_size_decodes(codec, oracle):
    """The reference checks a block's sample count against the caller's buffer only (linne_decoder.c:632-635),
    so a stream whose header names a SMALLER block size than its blocks carry decodes all the same.  The
    kernels' shared-memory lines must be sized from the blocks, not from the header."""
    pcm = harness.synth_pcm(n=2048 * 3 + 300, channels=2, bits=16, seed=5)
    for preset in (0, 5):
        stream = bytearray(oracle.encode(pcm, preset=preset, block=2048))
        stream[24:28] = (1024).to_bytes(4, "big")                  # header.num_samples_per_block
        rc, out = codec.decode(bytes(stream), return_code=True)
        assert rc == OK
        assert np.array_equal(out, pcm), preset


def test_32_bit_pcm_is_refused(codec):
    """bits_per_sample + 1 bits of pre-emphasis state do not fit the format's 32-bit fields (bit_stream.h:317):
    the reference accepts the parameter and then shifts out of range; we refuse it up front."""
    skip_on_reference(codec, "accepts 32 bits per sample and corrupts the pre-emphasis field")
    L = codec.lib
    buf = np.zeros(64, np.uint8)
    h = valid_header(); h.bits_per_sample = 32
    assert L.LINNEEncoder_EncodeHeader(C.byref(h), u8(buf), 30) == INVALID_FORMAT
    h.bits_per_sample = 31
    assert L.LINNEEncoder_EncodeHeader(C.byref(h), u8(buf), 30) == OK
    enc = codec.make_encoder(2, 4096)
    try:
        prm = LINNEEncodeParameter(2, 32, 44100, 4096, 0, 1, 0, 0)
        assert L.LINNEEncoder_SetEncodeParameter(enc, C.byref(prm)) == INVALID_FORMAT
    finally:
        L.LINNEEncoder_Destroy(enc)

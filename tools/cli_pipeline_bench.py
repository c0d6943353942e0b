#!/usr/bin/env python
"""File pipeline of the batch command-line driver on real files.

Writes `--files` stereo 16-bit WAV files of `--seconds` each to a scratch directory (tmpfs when available), runs
`linne_b200_cli -e` and `-d` over the list with 1 and with N workers, checks that the decoded files equal the
inputs, and prints one JSON line with the tool's own timing summaries (`-s`: file read + transfers + kernels + file
write of every file, handle creation excluded)."""
import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    import harness
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--preset", type=int, default=0)
    ap.add_argument("--workers", type=int, nargs="+", default=[1, 2, 4])
    args = ap.parse_args()
    cli = os.path.join(ROOT, "linne_b200", "linne_b200_cli")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="lnb_cli_", dir=base)
    try:
        clip = harness.synth_pcm(seconds=10.0, sr=44100, channels=2, bits=16, seed=1)
        n = int(args.seconds * 44100)
        reps = (n + clip.shape[1] - 1) // clip.shape[1]
        wavs = []
        for i in range(args.files):
            pcm = np.tile(np.roll(clip, 977 * i, axis=1), (1, reps))[:, :n]
            data = np.ascontiguousarray(pcm.T).astype("<i2").tobytes()
            hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack(
                "<IHHIIHH", 16, 1, 2, 44100, 44100 * 4, 4, 16) + b"data" + struct.pack("<I", len(data))
            path = os.path.join(tmp, f"f{i:04d}.wav")
            with open(path, "wb") as f:
                f.write(hdr + data)
            wavs.append(path)
        out = {"files": args.files, "seconds_per_file": args.seconds, "preset": args.preset, "scratch": tmp, "runs": {}}
        for j in args.workers:
            enc_list, dec_list = os.path.join(tmp, f"enc{j}.txt"), os.path.join(tmp, f"dec{j}.txt")
            with open(enc_list, "w") as f:
                f.write("".join(f"{w} {w}.j{j}.lnn\n" for w in wavs))
            with open(dec_list, "w") as f:
                f.write("".join(f"{w}.j{j}.lnn {w}.j{j}.wav\n" for w in wavs))
            stats = {}
            for tag, cmd in (("encode", [cli, "-e", "-m", str(args.preset), "-j", str(j), "-s", "-L", enc_list]),
                             ("decode", [cli, "-d", "-j", str(j), "-s", "-L", dec_list])):
                res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
                if res.returncode != 0:
                    raise SystemExit(f"{cmd}: {res.stderr[-400:]}")
                stats[tag] = json.loads(res.stderr.strip().splitlines()[-1])
            same = all(open(w, "rb").read()[44:] == open(f"{w}.j{j}.wav", "rb").read()[44:] for w in wavs)
            ident = all(open(f"{w}.j{j}.lnn", "rb").read() == open(f"{w}.j{args.workers[0]}.lnn", "rb").read() for w in wavs)
            out["runs"][f"workers{j}"] = {"encode_msamples_s": stats["encode"]["msamples_per_s"], "encode_s": stats["encode"]["seconds"],
                                          "decode_msamples_s": stats["decode"]["msamples_per_s"], "decode_s": stats["decode"]["seconds"],
                                          "lossless": same, "streams_equal_first_run": ident}
        print(json.dumps(out))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

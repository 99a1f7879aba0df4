"""Copies the reference's audio assets that BASELINE.json's configs name into tests/golden/assets/ (run once in the
container that has /root/reference; the GPU box only sees the committed copies).

  YuaiLoop.wav     48 kHz / 24-bit / stereo, 288 000 frames   -> configs[0] at 44.1 kHz output (ratio >= 1 branch)
  bass.wav         44.1 kHz / 16-bit / mono + smpl loop        -> configs[0]'s "44.1 -> 48 kHz" branch
  pad-ambient.wav  48 kHz / float32 / stereo + smpl loop       -> configs[3] granular source

The files are data (test inputs), not reference source code; they are decoded at test time by the product's and the
oracle's own RIFF/WAVE readers (tests/test_wav_io.py checks both against a third numpy reading).
"""
import hashlib
import os
import shutil

SRC = "/root/reference/assets"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
FILES = ["YuaiLoop.wav", "bass.wav", "pad-ambient.wav"]

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        os.chmod(os.path.join(DST, f), 0o644)
        print(f, hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest())

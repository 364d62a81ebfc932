#!/usr/bin/env python
"""Write plans/*.ini: the VFO plans of the reference's sample configurations reduced to the
keys the channelizer hot path reads (mainwindow.cpp:29-223), comments and device/GUI keys
dropped. Run once in the authoring container (needs /root/reference); the GPU box has only
plans/. The reduced files build the identical plan -- tests/test_plan.py checks that where
the originals are available."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.plan import read_ini  # noqa: E402

NAMES = {"sdr_25E": "25E", "sdr_98W": "98W", "sdr_54W_all": "54W_all", "sdr_54W_288K": "54W_288K",
         "CBAND_143E": "CBAND_143E"}
TOP = ["sample_rate", "center_frequency", "zmq_address", "correct_dc_bias", "mix_offset"]
MAIN = ["frequency", "out_rate", "zmq_address", "zmq_topic", "compress_scale"]
SUB = ["frequency", "gain", "data_rate", "out_rate", "filter_bandwidth", "topic"]


def main(src="/root/reference/sample_ini", dst=None):
    dst = dst or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "plans")
    os.makedirs(dst, exist_ok=True)
    for name, short in NAMES.items():
        kv = read_ini(os.path.join(src, name + ".ini"))
        lines = ["; plan %s: hot-path keys of the reference's sample_ini/%s.ini" % (short, name)]
        for k in TOP:
            if k in kv:
                lines.append("%s=%s" % (k, kv[k]))
        for sec, keys in (("main_vfos", MAIN), ("vfos", SUB)):
            n = int(kv.get(sec + "/size", "0"))
            lines += ["", "[%s]" % sec, "size=%d" % n]
            for i in range(1, n + 1):
                for k in keys:
                    full = "%s/%d/%s" % (sec, i, k)
                    if full in kv:
                        lines.append("%d\\%s=%s" % (i, k, kv[full]))
        with open(os.path.join(dst, short + ".ini"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("wrote", short)


if __name__ == "__main__":
    main(*sys.argv[1:])

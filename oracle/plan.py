"""Oracle-side ini -> VFO plan (test infrastructure; the product has its own, in C++).

Restates the QSettings(IniFormat) reading and the plan arithmetic of
/root/reference/mainwindow.cpp:27-239 with Python ints/floats. Used only to drive
oracle/sdr_oracle.c and to cross-check the product's plan compiler.
"""
import math
import os

SUPPORTED_RATES = (288000, 1536000, 1920000)           # mainwindow.h:29


def read_ini(path):
    """QSettings IniFormat subset: returns {"group/key": "value"}; ';' starts a
    comment, '#' does not (such lines become harmless junk keys), '\\' in a key
    is the array separator, later duplicates win, top-level keys have no group."""
    kv = {}
    group = ""
    with open(path, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line or line[0] == ";":
                continue
            if line[0] == "[":
                group = line[1:line.index("]")].strip() if "]" in line else line[1:].strip()
                if group == "General":
                    group = ""
                continue
            if "=" not in line:
                continue
            k, v = line.split("=", 1)
            k = k.strip().replace("\\", "/")
            v = v.strip()
            if len(v) >= 2 and v[0] == '"' and v[-1] == '"':
                v = v[1:-1]
            kv[(group + "/" + k) if group else k] = v
    return kv


def _to_int(s):
    """QString::toInt(): whole string must be a base-10 int32, else 0."""
    try:
        v = int(s.strip(), 10)
    except (ValueError, AttributeError):
        return 0
    return v if -2**31 <= v < 2**31 else 0


def _to_float(s):
    try:
        return float(s)
    except (ValueError, TypeError):
        return 0.0


def build_plan(path):
    kv = read_ini(path)
    g = lambda k: kv.get(k, "")
    Fs = _to_int(g("sample_rate"))
    if Fs not in SUPPORTED_RATES:
        raise ValueError("sample_rate %r not supported" % Fs)
    center = _to_int(g("center_frequency"))
    mix_offset = _to_int(g("mix_offset"))
    bufsplit = 4                                        # mainwindow.cpp:67-80
    if ((2 * Fs) // 4) % 512 > 0:
        buflen = (2 * Fs) // 5
        bufsplit = 5
    else:
        buflen = (2 * Fs) // 4
    plan = {
        "name": os.path.splitext(os.path.basename(path))[0],
        "Fs": Fs, "center": center, "bufsplit": bufsplit, "buflen": buflen,
        "block": buflen // 2, "dc": g("correct_dc_bias") == "1",
        "zmq_address": g("zmq_address"), "mains": [], "subs": [],
    }
    for i in range(_to_int(g("main_vfos/size"))):       # mainwindow.cpp:98-140
        p = "main_vfos/%d/" % (i + 1)
        freq = _to_int(g(p + "frequency"))
        out_rate = _to_int(g(p + "out_rate"))
        decim = 0 if Fs // out_rate == 1 else int(math.log2(Fs // out_rate))
        scalecomp = _to_int(g(p + "compress_scale"))    # mainwindow.cpp:112-118; vfo.cpp:24 default 1
        addr, topic = g(p + "zmq_address"), g(p + "zmq_topic")
        if not (addr and topic):                        # mainwindow.cpp:120-126: both or neither
            addr, topic = "", ""
        plan["mains"].append({
            "freq": freq, "mixer": float(center - freq), "decim": decim,
            "out_rate": int(Fs / (2 ** decim)), "samples_per_buffer": buflen // 2,
            "topic": topic, "zmq_address": addr, "scalecomp": scalecomp if scalecomp > 0 else 1,
        })
    for i in range(_to_int(g("vfos/size"))):            # mainwindow.cpp:141-235
        p = "vfos/%d/" % (i + 1)
        freq = _to_int(g(p + "frequency")) + mix_offset
        data_rate = _to_int(g(p + "data_rate"))
        out_rate = _to_int(g(p + "out_rate"))
        if out_rate == 0 and data_rate > 0:
            out_rate = {600: 12000, 1200: 24000}.get(data_rate, 48000)
        filterbw = _to_int(g(p + "filter_bandwidth"))
        main_freq, main_out, main_idx = 0, Fs, 0
        for a, m in enumerate(plan["mains"]):
            diff = int(abs((center - m["mixer"]) - freq))
            if diff < m["out_rate"]:
                main_idx, main_freq, main_out = a, int(m["mixer"]), m["out_rate"]
                break
        late = 0
        if main_out // 48000 == 5:
            decim = int(math.log2(main_out // (5 * out_rate)))
            late = 5
        elif main_out // 48000 == 6:
            decim = int(math.log2(main_out // (6 * out_rate)))
            late = 6
        else:
            decim = int(math.log2(Fs // out_rate)) - int(math.log2(Fs // main_out))
        import numpy as np
        gain = float(np.float32(np.float32(_to_float(g(p + "gain"))) / np.float32(100)))
        spb = main_out // bufsplit
        rate = int(main_out / (2 ** decim))
        samples_out = int(spb / (2 ** decim))
        if late:
            rate //= late
            samples_out //= late
        plan["subs"].append({
            "topic": g(p + "topic"), "freq": freq, "main": main_idx, "Fs": main_out,
            "mixer": float((center - main_freq) - freq), "decim": decim, "late": late,
            "filterbw": filterbw, "gain": gain, "samples_per_buffer": spb,
            "out_rate": rate, "samples_out": samples_out, "data_rate": data_rate,
        })
    return plan

"""Deterministic synthetic RTL-SDR uint8 IQ.

One counter-based generator (splitmix64 -> Box-Muller, all in numpy integer /
float64 arithmetic, no library RNG) so that the very same bytes can be handed to
the reference oracle and to the GPU path on any box. Per stream `s` the seed is
0x5D2B200 + s.
"""
import numpy as np

SEED0 = 0x5D2B200
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _uniform(seed, lane, idx):
    """float64 in (0,1) from (seed, lane, index)."""
    with np.errstate(over="ignore"):
        k = _splitmix64(np.uint64(seed) ^ (np.uint64(lane) * np.uint64(0xD1342543DE82EF95)))
        z = _splitmix64(idx.astype(np.uint64) ^ k)
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _bits(seed, lane, idx):
    with np.errstate(over="ignore"):
        k = _splitmix64(np.uint64(seed) ^ (np.uint64(lane) * np.uint64(0xD1342543DE82EF95)))
        z = _splitmix64(idx.astype(np.uint64) ^ k)
    return (z >> np.uint64(63)).astype(np.float64)


def carriers_for_plan(center, subs, offset_hz=1500.0, amp=1.0):
    """One BPSK-like carrier per sub-VFO, 1.5 kHz above its dial frequency (so it
    lands in the USB audio band), plus one un-channelised interferer at +250 kHz."""
    cs = [(float(s["freq"]) + offset_hz - center, amp, float(s.get("data_rate") or 600)) for s in subs]
    cs.append((250000.0, 6.0, 0.0))
    return cs


def make_iq(fs, n_samples, carriers, stream=0, start=0, sigma=2.5, dc=127.5, level=1.0, chunk=1 << 20):
    """uint8 interleaved I,Q for complex samples [start, start+n_samples) of `stream`.

    carriers: list of (offset_hz_from_center, amplitude_lsb, symbol_rate_or_0).
    level scales noise and carriers together; plans whose sub-VFOs run the gain-2
    /5 or /6 decimating FIR (vfo.cpp:75-87) need level=0.5 to stay inside int16,
    where the reference's unsaturated conversion (vfo.cpp:328,364) is defined."""
    seed = SEED0 + int(stream)
    out = np.empty(2 * n_samples, dtype=np.uint8)
    fs_i = int(fs)
    for c0 in range(0, n_samples, chunk):
        n = np.arange(start + c0, start + min(c0 + chunk, n_samples), dtype=np.int64)
        u1 = _uniform(seed, 1, n)
        u2 = _uniform(seed, 2, n)
        r = (sigma * level) * np.sqrt(-2.0 * np.log(u1))
        xi = dc + r * np.cos(2 * np.pi * u2)
        xq = dc + r * np.sin(2 * np.pi * u2)
        for lane, (f, a, baud) in enumerate(carriers):
            fi = int(round(f))
            ph = 2 * np.pi * (((n * fi) % fs_i).astype(np.float64) / fs_i)
            if baud > 0:
                sym = (n * int(baud)) // fs_i
                ph = ph + np.pi * _bits(seed, 100 + lane, sym)
            xi += (a * level) * np.cos(ph)
            xq += (a * level) * np.sin(ph)
        seg = out[2 * c0: 2 * (c0 + n.size)]
        seg[0::2] = np.clip(np.rint(xi), 0, 255).astype(np.uint8)
        seg[1::2] = np.clip(np.rint(xq), 0, 255).astype(np.uint8)
    return out

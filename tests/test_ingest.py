"""Ingest front ends (host side, no device, no socket): the rtl_tcp framing of sdrj.cpp:31-74,125-188
and the 20-buffer callback ring of jonti/sdr.cpp:100-184, through the C ABI."""
import ctypes as C
import threading

import numpy as np

from sdrreceiver_b200 import binding as B


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_rtltcp_header_and_block_framing():
    L = B.lib()
    h = C.c_void_p()
    assert L.sdrb_rtltcp_create(288000, 0, C.byref(h)) == 0
    blk = L.sdrb_rtltcp_block_bytes(h)
    assert blk == (288000 // 4) * 2                                   # sdrj.cpp:46
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, size=3 * blk + 1000, dtype=np.uint8)
    header = b"RTL0" + (5).to_bytes(4, "big") + (29).to_bytes(4, "big")  # tuner type R820T, 29 gain steps
    stream = np.frombuffer(header + data.tobytes(), dtype=np.uint8)
    tt, gc = C.c_uint32(), C.c_uint32()
    assert L.sdrb_rtltcp_header(h, C.byref(tt), C.byref(gc)) == 0
    at, sizes, k = 0, [1, 2, 5, 3, 1, 4096, 70000, 13, 200000], 0   # the socket delivers arbitrary pieces
    ready = 0
    while at < stream.size:
        n = min(sizes[k % len(sizes)], stream.size - at); k += 1
        piece = np.ascontiguousarray(stream[at:at + n])
        ready = L.sdrb_rtltcp_feed(h, _p(piece), n)
        assert ready >= 0
        at += n
    assert L.sdrb_rtltcp_header(h, C.byref(tt), C.byref(gc)) == 1 and (tt.value, gc.value) == (5, 29)
    assert ready == 3
    out = np.zeros(blk, np.uint8)
    for b in range(3):
        assert L.sdrb_rtltcp_pop(h, _p(out)) == 1
        assert np.array_equal(out, data[b * blk:(b + 1) * blk])
    assert L.sdrb_rtltcp_pop(h, _p(out)) == 0                          # 1000 bytes wait for the rest of block 4
    rest = np.ascontiguousarray(rng.integers(0, 256, size=blk - 1000, dtype=np.uint8))
    assert L.sdrb_rtltcp_feed(h, _p(rest), rest.size) == 1
    assert L.sdrb_rtltcp_pop(h, _p(out)) == 1
    assert np.array_equal(out[:1000], data[3 * blk:]) and np.array_equal(out[1000:], rest)
    L.sdrb_rtltcp_destroy(h)
    # a source without the dongle header (a recording), and a caller-chosen block (5 callbacks/s plan)
    assert L.sdrb_rtltcp_create(288000, 2 * 57600, C.byref(h)) == 0
    assert L.sdrb_rtltcp_block_bytes(h) == 115200
    raw = np.ascontiguousarray(data[:115200 + 7])
    assert L.sdrb_rtltcp_feed(h, _p(raw), raw.size) == 1
    assert L.sdrb_rtltcp_header(h, None, None) == 0
    out2 = np.zeros(115200, np.uint8)
    assert L.sdrb_rtltcp_pop(h, _p(out2)) == 1 and np.array_equal(out2, raw[:115200])
    L.sdrb_rtltcp_destroy(h)


def test_rtltcp_commands_are_big_endian():
    L = B.lib()
    out = np.zeros(5, np.uint8)
    L.sdrb_rtltcp_command(0x01, 1545600000, _p(out))                   # CMD_SET_FREQ, sdrj.cpp:176-182
    assert bytes(out) == b"\x01" + (1545600000).to_bytes(4, "big")
    seq = np.zeros(25, np.uint8)
    assert L.sdrb_rtltcp_start_sequence(1536000, 1545600000, 14, _p(seq)) == 25
    want = (b"\x08" + (0).to_bytes(4, "big") + b"\x03" + (1).to_bytes(4, "big") + b"\x0d" + (14).to_bytes(4, "big") +
            b"\x02" + (1536000).to_bytes(4, "big") + b"\x01" + (1545600000).to_bytes(4, "big"))   # sdrj.cpp:56-66
    assert bytes(seq) == want


def test_ring_is_fifo_and_drops_when_full():
    L = B.lib()
    r = C.c_void_p()
    assert L.sdrb_ring_create(4096, 0, 0, C.byref(r)) == 0            # 0 buffers = N_BUFFERS = 20 (jonti/sdr.h:83)
    bufs = [np.full(4096, k, np.uint8) for k in range(23)]
    res = [L.sdrb_ring_push(r, _p(b), 4096 if k != 5 else 100) for k, b in enumerate(bufs)]
    assert res == [1] * 20 + [0] * 3                                   # the 21st..23rd are dropped (sdr.cpp:103-110)
    pushed, dropped, used = C.c_uint64(), C.c_uint64(), C.c_int()
    L.sdrb_ring_stats(r, C.byref(pushed), C.byref(dropped), C.byref(used))
    assert (pushed.value, dropped.value, used.value) == (20, 3, 20)
    ptr, ln = C.c_void_p(), C.c_uint32()
    for k in range(20):
        assert L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == 1
        got = np.ctypeslib.as_array((C.c_uint8 * ln.value).from_address(ptr.value))
        assert ln.value == (4096 if k != 5 else 100) and np.all(got == k)   # buffers_size_valid travels with the slot
        assert L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == -1      # one buffer on loan at a time
        assert L.sdrb_ring_release(r) == 0
    assert L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 10) == 0          # empty: times out
    assert L.sdrb_ring_push(r, _p(bufs[0]), 5000) == -1                    # longer than a buffer
    L.sdrb_ring_destroy(r)


def test_ring_between_two_threads_and_cancel():
    L = B.lib()
    r = C.c_void_p()
    assert L.sdrb_ring_create(1024, 4, 0, C.byref(r)) == 0
    n, seen = 300, []

    def consumer():
        ptr, ln = C.c_void_p(), C.c_uint32()
        while L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), -1) == 1:
            a = np.ctypeslib.as_array((C.c_uint8 * ln.value).from_address(ptr.value))
            seen.append(int(a[0]) | int(a[1]) << 8)
            L.sdrb_ring_release(r)

    th = threading.Thread(target=consumer)
    th.start()
    sent = []
    for k in range(n):
        b = np.zeros(1024, np.uint8); b[0], b[1] = k & 255, k >> 8
        if L.sdrb_ring_push(r, _p(b), 1024) == 1:
            sent.append(k)
    while True:                                                         # drain, then stop the dispatcher
        used = C.c_int()
        L.sdrb_ring_stats(r, None, None, C.byref(used))
        if used.value == 0:
            break
    L.sdrb_ring_cancel(r)
    th.join(timeout=10)
    assert not th.is_alive()
    assert seen == sent and len(sent) >= 4                              # order kept; what was dropped never shows up
    dropped = C.c_uint64()
    L.sdrb_ring_stats(r, None, C.byref(dropped), None)
    assert dropped.value == n - len(sent)
    L.sdrb_ring_destroy(r)


import pytest  # noqa: E402
from conftest import plan_path  # noqa: E402
from oracle import oracle as O, plan as OP  # noqa: E402
from sdrreceiver_b200 import synth  # noqa: E402

needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (needs /root/reference once)")


def _plan_input(name, n_blocks, level=1.0):
    op = OP.build_plan(plan_path(name))
    return op, synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]), level=level)


@needs_ref
def test_reference_callback_ring_is_the_oracle_of_sdrb_ring():
    """The reference's own librtlsdr entry point as the oracle of row f2: bytes pushed through sdr::rtlsdr_callback and
    drained by sdr::demod_dispatcher (jonti/sdr.cpp:100-184; the harness plays both threads) give exactly what the
    direct feed gives; with 23 callbacks queued before the dispatcher runs the reference keeps the first 20 and drops
    three -- and sdrb_ring does the same with the same buffers."""
    L = B.lib()
    op, iq = _plan_input("54W_288K", 23, level=0.5)
    blk = op["block"] * 2
    ini = plan_path("54W_288K")
    direct6, frames6, _ = O.run_ref(ini, iq, blocks=6)
    via2, framesv, _, info = O.run_ref(ini, iq, blocks=6, via="callback:2")
    assert info["ingest"] == {"callbacks": 6, "burst": 2} and framesv == frames6
    for k in direct6:
        assert np.array_equal(via2[k], direct6[k]), k
    # overflow: 23 callbacks arrive while the dispatcher sleeps
    direct20, _, _ = O.run_ref(ini, iq, blocks=20)
    via23, _, _, _ = O.run_ref(ini, iq, blocks=23, via="callback:23")
    for k in direct20:
        assert np.array_equal(via23[k], direct20[k]), k                 # the 21st..23rd never reach demodData
    r = C.c_void_p()
    assert L.sdrb_ring_create(blk, 0, 0, C.byref(r)) == 0
    res = [L.sdrb_ring_push(r, _p(np.ascontiguousarray(iq[k * blk:(k + 1) * blk])), blk) for k in range(23)]
    assert res == [1] * 20 + [0] * 3
    ptr, ln = C.c_void_p(), C.c_uint32()
    for k in range(20):
        assert L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == 1 and ln.value == blk
        got = np.ctypeslib.as_array((C.c_uint8 * blk).from_address(ptr.value))
        assert np.array_equal(got, iq[k * blk:(k + 1) * blk])           # what the reference processed, in its order
        L.sdrb_ring_release(r)
    assert L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == 0
    L.sdrb_ring_destroy(r)


@needs_ref
def test_reference_rtl_tcp_client_is_the_oracle_of_sdrb_rtltcp():
    """sdrj::start_tcp_rtl / sdrj::readyRead (sdrj.cpp:31-74,125-166) driven through an in-memory socket: the command
    bytes the client sends, the 12-byte greeting, and the block framing of an arbitrarily chopped stream -- each against
    sdrb_rtltcp_*. The audio the reference produces from the socket equals the direct feed of the same bytes."""
    L = B.lib()
    op, iq = _plan_input("CBAND_143E", 3)
    ini = plan_path("CBAND_143E")
    blk = op["block"] * 2
    sizes = [100000, 250000, 7, 300000, 65536, 511]
    direct, frames, _ = O.run_ref(ini, iq)
    via, framesv, _, info = O.run_ref(ini, iq, via="rtltcp:" + ",".join(map(str, sizes)))
    assert info["ingest"]["left_in_socket"] == 0 and info["ingest"]["block_bytes"] == blk == (op["Fs"] // 4) * 2
    assert framesv == frames
    for k in direct:
        assert np.array_equal(via[k], direct[k]), k
    seq = np.zeros(25, np.uint8)
    assert L.sdrb_rtltcp_start_sequence(op["Fs"], op["center"], 14, _p(seq)) == 25
    assert bytes(seq) == info["tcp_tx"]                                 # AGC off, manual gain, gain index, rate, frequency
    # our framer on the same arrivals (greeting first, on its own -- the reference only recognises it that way)
    h = C.c_void_p()
    assert L.sdrb_rtltcp_create(op["Fs"], 0, C.byref(h)) == 0 and L.sdrb_rtltcp_block_bytes(h) == blk
    hello = np.frombuffer(b"RTL0" + (5).to_bytes(4, "big") + (29).to_bytes(4, "big"), np.uint8)
    assert L.sdrb_rtltcp_feed(h, _p(np.ascontiguousarray(hello)), 12) == 0
    at, k, out, popped = 0, 0, np.zeros(blk, np.uint8), 0
    while at < iq.size:
        n = min(sizes[k % len(sizes)], iq.size - at); k += 1
        L.sdrb_rtltcp_feed(h, _p(np.ascontiguousarray(iq[at:at + n])), n)
        at += n
        while L.sdrb_rtltcp_pop(h, _p(out)) == 1:
            assert np.array_equal(out, iq[popped * blk:(popped + 1) * blk])
            popped += 1
    assert popped == 3 and k == info["ingest"]["arrivals"]
    tt, gc = C.c_uint32(), C.c_uint32()
    assert L.sdrb_rtltcp_header(h, C.byref(tt), C.byref(gc)) == 1 and (tt.value, gc.value) == (5, 29)
    L.sdrb_rtltcp_destroy(h)


@pytest.mark.gpu
@needs_ref
def test_socket_to_audio_against_the_reference_rtl_tcp_client():
    """Row f2 end to end: the same chopped rtl_tcp stream through sdrb_rtltcp -> pinned sdrb_ring -> sdrb_bank_process_host
    and through the reference's own client (sdrj::readyRead -> demodData -> vfo tree): int16 within +-1 LSB."""
    L = B.lib()
    op, iq = _plan_input("CBAND_143E", 3)
    ini = plan_path("CBAND_143E")
    plan = B.Plan(ini)
    blk = plan.block * 2
    sizes = [100000, 250000, 7, 300000, 65536, 511]
    want, _, _, _ = O.run_ref(ini, iq, via="rtltcp:" + ",".join(map(str, sizes)))
    f, r = C.c_void_p(), C.c_void_p()
    assert L.sdrb_rtltcp_create(op["Fs"], 0, C.byref(f)) == 0
    assert L.sdrb_ring_create(blk, 0, 1, C.byref(r)) == 0
    stream = np.frombuffer(b"RTL0" + bytes(8) + iq.tobytes(), dtype=np.uint8)
    bank = B.Bank(plan, 1, 1)
    tmp, got, at, k = np.zeros(blk, np.uint8), [], 0, 0
    while at < stream.size:
        n = 12 if at == 0 else min(sizes[k % len(sizes)], stream.size - at)
        k += at != 0
        piece = np.ascontiguousarray(stream[at:at + n]); at += n
        L.sdrb_rtltcp_feed(f, _p(piece), n)
        while L.sdrb_rtltcp_pop(f, _p(tmp)) == 1:
            assert L.sdrb_ring_push(r, _p(tmp), blk) == 1
        ptr, ln = C.c_void_p(), C.c_uint32()
        while L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == 1:
            pcm = np.zeros(plan.pcm_per_block, np.int16)
            bank.process_host(ptr, blk, 1, _p(pcm))
            got.append(pcm)
            L.sdrb_ring_release(r)
    assert len(got) == 3
    mine = B.split_pcm(plan, np.stack(got))
    for s in op["subs"]:
        d = np.abs(mine[s["topic"]].astype(np.int32) - want[s["topic"]].astype(np.int32)).max()
        assert d <= 1, (s["topic"], d)
    bank.close()
    L.sdrb_ring_destroy(r); L.sdrb_rtltcp_destroy(f)


@pytest.mark.gpu
def test_rtltcp_stream_through_pinned_ring_into_the_bank():
    """Socket bytes -> framer -> pinned ring -> sdrb_bank_process_host straight out of the ring buffer:
    identical to feeding the same callbacks directly."""
    from conftest import plan_path
    from oracle import plan as OP
    from sdrreceiver_b200 import synth
    L = B.lib()
    op = OP.build_plan(plan_path("54W_288K")); plan = B.Plan(plan_path("54W_288K"))
    n_blocks, blk = 3, plan.block * 2
    iq = synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]), level=0.5)
    direct = B.Bank(plan, 1, 1)
    want = [direct.process_numpy(iq[None, k * blk:(k + 1) * blk], 1)[0][0, 0].copy() for k in range(n_blocks)]
    direct.close()
    f, r = C.c_void_p(), C.c_void_p()
    assert L.sdrb_rtltcp_create(op["Fs"], blk, C.byref(f)) == 0          # this plan runs 5 callbacks/s: own block size
    assert L.sdrb_ring_create(blk, 0, 1, C.byref(r)) == 0                # pinned buffers
    stream = np.frombuffer(b"RTL0" + bytes(8) + iq.tobytes(), dtype=np.uint8)
    bank = B.Bank(plan, 1, 1)
    tmp, got, at = np.zeros(blk, np.uint8), [], 0
    while at < stream.size:
        n = min(50000, stream.size - at)
        piece = np.ascontiguousarray(stream[at:at + n]); at += n
        L.sdrb_rtltcp_feed(f, _p(piece), n)
        while L.sdrb_rtltcp_pop(f, _p(tmp)) == 1:
            assert L.sdrb_ring_push(r, _p(tmp), blk) == 1
        ptr, ln = C.c_void_p(), C.c_uint32()
        while L.sdrb_ring_pop(r, C.byref(ptr), C.byref(ln), 0) == 1:
            pcm = np.zeros(plan.pcm_per_block, np.int16)
            bank.process_host(ptr, blk, 1, _p(pcm))
            got.append(pcm)
            L.sdrb_ring_release(r)
    assert len(got) == n_blocks
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    bank.close()
    L.sdrb_ring_destroy(r); L.sdrb_rtltcp_destroy(f)

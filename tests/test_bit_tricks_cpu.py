"""CPU restatements of three small constructions in csrc/ (the compiled kernels are checked bit for bit against their
plain siblings on the GPU: test_kernels_gpu.py::test_bn_fast_paths_are_bit_identical / test_tiled_stem_maxpool_is_bit_identical;
these tests pin the ARITHMETIC the constructions rely on, for every input, without a device).

* vec.cuh relu_lane_mask: ReLU bit mask (bit j = channel j of an 8-channel vector) -> 16-bit lane masks of the four
  packed bf16 words with one multiply and one prmt.b32 in sign-replicate mode;
* norm_act.cu bn_relu_maxpool_tiled_kernel<true>: the max-pool scan runs on sign-flipped bf16 words with -inf (0xFF80)
  outside the image, strict > keeps the first maximum;
* norm_act.cu bn_bwd_apply_async_kernel / bn_bwd_reduce_async_kernel: the slot / commit-group bookkeeping of the
  thread-private cp.async ring, for any number of vectors per thread."""
import numpy as np


def prmt_default(a, b, sel):
    """PTX prmt.b32 default mode: result byte i = byte (sel nibble i & 7) of {b, a}; nibble bit 3 replicates its sign."""
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        byte = src[nib & 7]
        if nib & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def relu_lane_mask(m8, k):
    t = (m8 * ((1 << (15 - 2 * k)) + (1 << (30 - 2 * k)))) & 0xFFFFFFFF
    return prmt_default(t, 0, 0xBB99)


def test_relu_lane_mask_for_every_mask_byte():
    for m8 in range(256):
        for k in range(4):
            want = (0x0000FFFF if (m8 >> (2 * k)) & 1 else 0) | (0xFFFF0000 if (m8 >> (2 * k + 1)) & 1 else 0)
            assert relu_lane_mask(m8, k) == want, (m8, k)


def test_masked_words_equal_unpack_select_pack():
    """dz word & lane mask == pack(select(unpack(dz))) for bf16 bit patterns including -0.0, inf and subnormals."""
    rng = np.random.default_rng(0)
    words = rng.integers(0, 2**32, size=4096, dtype=np.uint64).astype(np.uint32)
    words[:4] = [0x80008000, 0x7F80FF80, 0x00010001, 0x80000000]
    for w in words[:512]:
        for m2 in range(4):  # the two mask bits of this word
            lo = int(w) & 0xFFFF if m2 & 1 else 0      # masked lane -> +0.0, as `dz = 0.f` packs
            hi = int(w) >> 16 if m2 & 2 else 0
            mask = (0xFFFF if m2 & 1 else 0) | (0xFFFF0000 if m2 & 2 else 0)
            assert (int(w) & mask) == (lo | (hi << 16))


def bf16_key(h):
    """Total order of bf16 bit patterns as floats (no NaNs): value of the upper half of an fp32."""
    return np.array([h << 16], dtype=np.uint32).view(np.float32)[0]


def test_flipped_domain_scan_equals_first_maximum_of_the_affine_map():
    """max over a 3x3 window of relu(sc * x + sh) with torch's tie rule (first maximum in row-major order) equals: flip
    the sign bit where sc < 0, scan with strict >, treat out-of-image taps as -inf, un-flip the winner."""
    rng = np.random.default_rng(1)
    vals = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 2.0, -2.0, 3.5, -0.25], dtype=np.float32)
    for trial in range(400):
        x = rng.choice(vals, size=9)
        valid = rng.random(9) > 0.25
        valid[4] = True  # the window centre is always inside the image
        for sc in (0.7, -0.7):
            sh = 0.1
            xb = (x.view(np.uint32) >> 16).astype(np.uint32)  # exact: the values are bf16-representable
            flip = 0x8000 if sc < 0 else 0
            best, idx = 0xFF80, 0
            for t in range(9):
                h = int(xb[t]) ^ flip if valid[t] else 0xFF80
                if bf16_key(h) > bf16_key(best):
                    best, idx = h, t
            got = max(sc * float(bf16_key(best ^ flip)) + sh, 0.0)
            # reference: affine map first, then the first maximum over the valid taps
            a = [sc * float(x[t]) + sh if valid[t] else -np.inf for t in range(9)]
            want_t = int(np.argmax(a))  # first maximum
            assert abs(got - max(a[want_t], 0.0)) < 1e-6
            assert idx == want_t  # same tap as torch's tie rule (+0.0 and -0.0 tie in both orders)


# ------------------------------------------------------------------------------------------------------------------
# The thread-private cp.async ring of bn_bwd_apply_async_kernel / bn_bwd_reduce_async_kernel (norm_act.cu): one thread's
# bookkeeping, restated.  Vector v of the thread travels in commit group v; iteration v issues the copy of vector
# v + S - 1 into slot (v + S - 1) % S, commits, waits until at most S - 1 groups are pending and consumes slot v % S.
def _ring_thread(n_vectors, S, land_early):
    """Simulates the loop for a thread that owns n_vectors vectors.  land_early: copies land the moment they are issued
    (catches a slot overwritten before it was consumed); otherwise they land only when wait_group forces them (catches
    a slot consumed before its copy was guaranteed).  Returns the consumed vector ids in order."""
    slots = [None] * S              # what is VISIBLE in shared memory
    pending = []                    # committed groups not yet known complete: lists of (slot, vector)
    consumed, live = [], set()      # live: vectors copied (or in flight) and not yet consumed

    def issue(group, slot, v):
        assert all(x != slot for g in pending for x, _ in g), "slot has a copy in flight"
        assert slots[slot] is None or slots[slot] not in live, "overwrites a vector that was not consumed yet"
        live.add(v)
        if land_early:
            slots[slot] = v
            group.append((slot, None))
        else:
            group.append((slot, v))

    def wait(max_pending):
        while len(pending) > max_pending:
            for slot, v in pending.pop(0):
                if v is not None:
                    slots[slot] = v

    for s in range(S - 1):          # prologue
        g = []
        if s < n_vectors:
            issue(g, s, s)
        pending.append(g)
    v = 0
    while v < n_vectors:            # the kernel's two nested loops, flattened
        g = []
        if v + S - 1 < n_vectors:
            issue(g, (v + S - 1) % S, v + S - 1)
        pending.append(g)
        wait(S - 1)
        assert slots[v % S] == v, (v, slots)
        consumed.append(v)
        live.discard(v)
        v += 1
    assert not live
    return consumed


def test_cp_async_ring_bookkeeping():
    for S in (2, 3, 6, 8):
        for n in list(range(0, 3 * S + 2)) + [57, 123]:
            for early in (False, True):
                assert _ring_thread(n, S, early) == list(range(n)), (S, n, early)

"""TEST INFRASTRUCTURE ONLY — CPU oracle (Python twin) of the Fiat–Shamir transcript.

Follows joltworks/src/transcripts/blake2b.rs:11-258 (Blake2bTranscript).  Blake2b-256 itself
is RFC 7693 (`blake2 0.10.6` in the reference, `hashlib.blake2b(digest_size=32)` here; the
C++ host re-implements RFC 7693 and is checked against hashlib).
Parity unpinned at the byte level (no golden transcript states in the reference).
"""
from __future__ import annotations

import hashlib

from . import field as F


def _h(*chunks: bytes) -> bytes:
    h = hashlib.blake2b(digest_size=32)
    for c in chunks:
        h.update(c)
    return h.digest()


class Blake2bTranscript:
    def __init__(self, label: bytes):
        # blake2b.rs:81-100  new(label): H(label || zero-pad to 32)
        assert len(label) < 33
        self.state = _h(label + b"\0" * (32 - len(label)))
        self.n_rounds = 0
        self.state_history = [self.state]

    # blake2b.rs:31-37  hasher(): state(32) || 0^28 || n_rounds_be32
    def _prefix(self) -> bytes:
        return self.state + b"\0" * 28 + self.n_rounds.to_bytes(4, "big")

    def _update(self, new_state: bytes):
        self.state = new_state
        self.n_rounds += 1
        self.state_history.append(new_state)

    def append_message(self, msg: bytes):            # :109-122
        assert len(msg) < 33
        self._update(_h(self._prefix(), msg + b"\0" * (32 - len(msg))))

    def append_bytes(self, b: bytes):                # :124-128
        self._update(_h(self._prefix(), b))

    def append_u64(self, x: int):                    # :130-136
        self._update(_h(self._prefix(), b"\0" * 24 + x.to_bytes(8, "big")))

    def append_scalar(self, x: int):                 # :138-146  32-byte LE canonical, reversed
        self.append_bytes(F.fr_to_le_bytes(x)[::-1])

    def append_scalars(self, xs):                    # :158-164
        self.append_message(b"begin_append_vector")
        for x in xs:
            self.append_scalar(x)
        self.append_message(b"end_append_vector")

    def append_point(self, pt):                      # :166-187  pt = None (infinity) or (x, y) ints
        if pt is None:
            self.append_bytes(b"\0" * 64)
            return
        x, y = pt
        self._update(_h(self._prefix(), x.to_bytes(32, "big"), y.to_bytes(32, "big")))

    def append_points(self, pts):                    # :189-195
        self.append_message(b"begin_append_vector")
        for p in pts:
            self.append_point(p)
        self.append_message(b"end_append_vector")

    def append_serializable_bytes(self, le_bytes: bytes):   # :148-156 whole uncompressed string reversed
        self.append_bytes(le_bytes[::-1])

    def _challenge_bytes32(self) -> bytes:           # :58-63
        r = _h(self._prefix())
        self._update(r)
        return r

    def challenge_u128(self) -> int:                 # :197-202  reverse then from_be == LE of first 16 bytes
        return int.from_bytes(self._challenge_bytes32()[:16], "little")

    def challenge_scalar(self) -> int:               # :204-215  from_le_bytes_mod_order(reverse(first 16))
        return int.from_bytes(self._challenge_bytes32()[:16], "big") % F.P

    def challenge_vector(self, n):                   # :217-221
        return [self.challenge_scalar() for _ in range(n)]

    def challenge_scalar_powers(self, n):            # :224-231
        q = self.challenge_scalar()
        out = [1] * n
        for i in range(1, n):
            out[i] = out[i - 1] * q % F.P
        return out

    def challenge_scalar_optimized(self) -> int:     # :233-238  returns the *masked u128* (challenge id)
        return self.challenge_u128() & F.CHALLENGE_MASK

    def challenge_vector_optimized(self, n):         # :240-244
        return [self.challenge_scalar_optimized() for _ in range(n)]

"""FIFO replay buffer with episode markers -- API of smartstart/RLAgents/replay_buffer.py.

Behaviour (what may add, eviction, episode-number renormalisation, candidate sampling
with ``random.sample``, episodic path extraction) follows replay_buffer.py:27-215 and
is pinned by the reference's tests restated in tests/test_replay_buffer.py.

Re-design for the B200 path: next to the deque of experience
tuples the buffer keeps a contiguous float64 ring of ``s`` and ``s2`` rows, so
``get_all_states`` (KDE data set) and ``states_s2`` (KDE queries) are O(n) memcpys
instead of Python list comprehensions over a deque / ragged object arrays
(replay_buffer.py:102, smartexplorationcontinuous.py:272-273).
"""
from __future__ import annotations

import bisect
import pickle
import random
from collections import deque

import numpy as np


# --------------------------------------------------------------------------- candidate draw
_FAST_SAMPLE = None      # None: not probed yet; False: use the interpreter; else the library handle


_IN_PLACE = False        # the interpreter's generator state can be reached in place (probed)


def _state_in_place(rnd):
    """(address of the index, address of the 624 key words) of a ``random.Random``: CPython's
    RandomObject is {PyObject_HEAD; int index; uint32_t state[624]} (Modules/_randommodule.c).  An
    implementation detail, so it is only used after _probe_fast_sample has seen it agree with
    getstate() before and after a draw in this interpreter."""
    base = id(rnd) + object.__basicsize__
    return base, base + 4


def _sample_with(lib, rnd, first, n, k, in_place=False):
    """random.sample(range(first, first + n), k) of generator ``rnd`` through csrc/py_random.cu."""
    import ctypes as C
    out = np.empty(k, dtype=np.int64)
    if in_place:
        # no getstate() / setstate() round trip (2 x 625 Python ints: ~55 us of a ~0.3 ms draw)
        index_at, key_at = _state_in_place(rnd)
        rc = lib.ss_py_random_sample(C.c_void_p(key_at), C.byref(C.c_int.from_address(index_at)), int(first), int(n), int(k),
                                     out.ctypes.data_as(C.c_void_p))
        return out if rc == 0 else None
    version, internal, gauss = rnd.getstate()
    state = np.array(internal, dtype=np.uint32)                # 624 words + the index
    index = C.c_int(int(state[624]))
    rc = lib.ss_py_random_sample(state.ctypes.data_as(C.c_void_p), C.byref(index), int(first), int(n), int(k),
                                 out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        return None
    state[624] = index.value
    rnd.setstate((version, tuple(state.tolist()), gauss))
    return out


def _probe_fast_sample():
    """The C++ restatement is used only if it reproduces this interpreter's random.sample -- indices,
    order and the generator state afterwards -- on both of Random.sample's branches; the in-place
    route only if, on top of that, the memory behind the object is what getstate() reports."""
    global _FAST_SAMPLE, _IN_PLACE
    _FAST_SAMPLE = False
    _IN_PLACE = False
    try:
        from . import _lib
        lib = _lib.load()
        cases = ((7, 5000, 900), (3, 57, 25), (0, 1, 1), (11, 300000, 4000))
        for first, n, k in cases:
            a, b = random.Random(987654321), random.Random(987654321)
            want = a.sample(range(first, first + n), k)
            got = _sample_with(lib, b, first, n, k)
            if got is None or got.tolist() != want or a.getstate() != b.getstate():
                return
        _FAST_SAMPLE = lib
    except Exception:
        _FAST_SAMPLE = False
        return
    try:
        import ctypes as C
        import platform
        if platform.python_implementation() != "CPython":
            return
        for first, n, k in cases:
            a, b = random.Random(123456789), random.Random(123456789)
            a.random(), b.random()                                  # index away from 624
            index_at, key_at = _state_in_place(b)
            internal = b.getstate()[1]
            if C.c_int.from_address(index_at).value != internal[624] or \
                    list((C.c_uint32 * 624).from_address(key_at)) != list(internal[:624]):
                return
            want = a.sample(range(first, first + n), k)
            got = _sample_with(lib, b, first, n, k, in_place=True)
            if got is None or got.tolist() != want or a.getstate() != b.getstate() or a.random() != b.random():
                return
        _IN_PLACE = True
    except Exception:
        _IN_PLACE = False


def sample_range(first, stop, k):
    """np.array(random.sample(range(first, stop), k)) on the global generator (replay_buffer.py:152)."""
    if _FAST_SAMPLE is None:
        _probe_fast_sample()
    n = stop - first
    if _FAST_SAMPLE and k > 64 and 0 < n < 2 ** 31:
        rnd = getattr(random.sample, "__self__", None)       # the module-level generator random.sample is bound to
        if type(rnd) is random.Random:
            out = _sample_with(_FAST_SAMPLE, rnd, first, n, k, in_place=_IN_PLACE)
            if out is not None:
                return out
    return np.array(random.sample(range(first, stop), k))


class _StateRing:
    """Contiguous mirror of (s, s2) rows in buffer order."""

    def __init__(self, capacity, dim):
        self.capacity = capacity
        self.dim = dim
        self.alloc = min(capacity, 4096)
        self.s = np.empty((self.alloc, dim))
        self.s2 = np.empty((self.alloc, dim))
        self.head = 0        # physical row of buffer index 0
        self.count = 0
        self.pushes = 0      # total rows ever written: push p lives in physical row p % capacity

    def _grow(self):
        new_alloc = min(self.capacity, self.alloc * 2)
        for name in ("s", "s2"):
            old = getattr(self, name)
            new = np.empty((new_alloc, self.dim))
            new[:self.count] = old[:self.count]        # head is 0 until the ring is full
            setattr(self, name, new)
        self.alloc = new_alloc

    def push(self, s, s2):
        if self.count < self.capacity:
            if self.count == self.alloc:
                self._grow()
            row = self.count
            self.count += 1
        else:
            row = self.head
            self.head = (self.head + 1) % self.capacity
        self.s[row] = s
        self.s2[row] = s2
        self.pushes += 1

    def ordered(self, arr):
        if self.head == 0:
            return arr[:self.count]
        return np.concatenate([arr[self.head:self.count], arr[:self.head]], axis=0)

    def physical_rows(self, buffer_indices):
        """Physical rows of logical buffer indices 0 <= i < count (a fresh array)."""
        rows = np.asarray(buffer_indices, dtype=np.int64) + self.head
        if self.head:                                # wrapped ring: one conditional subtraction, no division
            rows[rows >= self.count] -= self.count
        return rows

    def rows(self, arr, buffer_indices):
        return arr[self.physical_rows(buffer_indices)]


class ReplayBuffer(object):
    """Stores (state, action, reward, terminal, new_state); right end = most recent.

    Only ``main_agent`` may add / mark episode starts (several agents share one buffer
    under SmartStart, replay_buffer.py:21-24).
    """

    def __init__(self, main_agent, max_buffer_size):
        self.main_agent = main_agent
        self.max_buffer_size = max_buffer_size
        # running count of adds, renormalised on eviction; episode starts hold counts
        self.next_episode_number = 0
        self.episode_starting_indices = deque()
        self.buffer = deque()
        self._ring = None
        self._ring_ok = True

    def set_main_agent(self, new_main_agent):
        self.main_agent = new_main_agent

    # ------------------------------------------------------------------ writes
    def add(self, observing_agent, s, a, r, t, s2):
        if observing_agent is not self.main_agent:
            return
        if len(self.buffer) >= self.max_buffer_size:
            starts = self.episode_starting_indices
            if starts and starts[0] == self.next_episode_number - self.max_buffer_size:
                # the oldest tracked episode loses its first step: stop tracking it and
                # shift all counts so the next tracked episode starts at 0
                starts.popleft()
                if starts:
                    shift = starts[0]
                    for i in range(len(starts)):
                        starts[i] -= shift
                    self.next_episode_number -= shift
                else:
                    self.next_episode_number = 0
            self.buffer.popleft()
        self.buffer.append((s, a, r, t, s2))
        self._mirror(s, s2)
        self.next_episode_number += 1

    def _mirror(self, s, s2):
        if not self._ring_ok:
            return
        try:
            fs = np.asarray(s, dtype=np.float64).reshape(-1)
            fs2 = np.asarray(s2, dtype=np.float64).reshape(-1)
            if self._ring is None:
                self._ring = _StateRing(self.max_buffer_size, fs.size)
            self._ring.push(fs, fs2)
        except (TypeError, ValueError):
            self._ring_ok = False      # non-numeric / ragged states: fall back to the deque
            self._ring = None

    def _rebuild_mirror(self):
        self._ring, self._ring_ok = None, True
        for step in self.buffer:
            self._mirror(step[0], step[4])

    def start_new_episode(self, observing_agent):
        if observing_agent is not self.main_agent:
            return
        starts = self.episode_starting_indices
        if starts and starts[-1] == self.next_episode_number:
            print("shouldn't call start episode twice in a row")
        else:
            starts.append(self.next_episode_number)

    def clear(self):
        self.buffer.clear()
        self.next_episode_number = 0
        self._ring = None
        self._ring_ok = True

    # ------------------------------------------------------------------ reads
    def size(self):
        return len(self.buffer)

    def __len__(self):
        return len(self.buffer)

    @staticmethod
    def _columns(batch):
        return tuple(np.array([step[c] for step in batch]) for c in range(5))

    def sample_batch(self, batch_size):
        if len(self.buffer) < batch_size:
            batch = random.sample(list(self.buffer), len(self.buffer))
        else:
            batch = random.sample(list(self.buffer), batch_size)
        return self._columns(batch)

    def all_batch(self):
        return self._columns(self.buffer)

    def get_all_states(self):
        """[n+1, d]: every stored s plus the newest s2 (the KDE data set)."""
        ring = self._ring
        if ring is not None and ring.count == len(self.buffer):
            s = ring.ordered(ring.s)
            last = ring.rows(ring.s2, [ring.count - 1])
            return np.concatenate([s, last], axis=0)
        return np.array([step[0] for step in self.buffer] + [self.buffer[-1][4]])

    def states_s2(self, buffer_indices):
        """s2 rows of the given buffer indices, [m, d] (the KDE queries)."""
        ring = self._ring
        if ring is not None and ring.count == len(self.buffer):
            return ring.rows(ring.s2, buffer_indices)
        return np.asarray([np.asarray(self.buffer[int(i)][4], dtype=np.float64)
                           for i in buffer_indices])

    def state_ring(self):
        """The contiguous (s, s2) ring when it mirrors the deque exactly, else None (the device
        mirror of Engine.select_start_mirror is built from it)."""
        ring = self._ring
        return ring if ring is not None and ring.count == len(self.buffer) else None

    def get_possible_smart_start_indices(self, n_ss):
        """Up to n_ss buffer indices sampled without replacement from the first fully
        stored episode onwards (the smart-start state is s2 of the step)."""
        if len(self.episode_starting_indices) == 0:
            return None
        first = self.episode_number_to_buffer_index(self.episode_starting_indices[0])
        n = min(n_ss, len(self.buffer) - first)
        return sample_range(first, len(self.buffer), n)

    def get_episodic_path_to_buffer_index(self, buffer_index):
        """States of the episode containing ``buffer_index`` from its start up to and
        including s2 of that step."""
        starts = self.episode_starting_indices
        if len(starts) == 0:
            raise ValueError(": (   -   no episodes have been recorded")
        count = self.buffer_index_to_episode_number(buffer_index)
        pos = bisect.bisect_right(starts, count)     # ascending; last start <= count wins
        start_count = starts[pos - 1] if pos > 0 else None
        first = self.episode_number_to_buffer_index(start_count)
        ring = self._ring
        if ring is not None and ring.count == len(self.buffer):
            # one gather (a private copy) and a list of its rows instead of one array construction per state
            rows = np.empty((buffer_index - first + 2, ring.dim))
            rows[:-1] = ring.rows(ring.s, np.arange(first, buffer_index + 1))
            rows[-1] = ring.s2[ring.physical_rows([buffer_index])[0]]
            return list(rows)
        steps = list(self.buffer)[first:buffer_index + 1]
        return [self.step_to_s(st) for st in steps] + [self.step_to_s2(self.buffer[buffer_index])]

    # ------------------------------------------------------------------ persistence
    def save(self):
        with open("replay_buffer.obj", "wb") as fh:
            pickle.dump(self.buffer, fh)
        print("the replay buffer was saved succesfully")

    def load(self):
        try:
            with open("replay_buffer.obj", "rb") as fh:
                self.buffer = pickle.load(fh)
            self.next_episode_number = len(self.buffer)
            self._rebuild_mirror()
            print("the replay buffer was loaded succesfully")
        except (OSError, pickle.UnpicklingError):
            print("there was no file to load")

    # ------------------------------------------------------------------ tuple accessors
    def step_to_s(self, step):
        return np.array(step[0])

    def step_to_a(self, step):
        return np.array(step[1])

    def step_to_r(self, step):
        return np.array(step[2])

    def step_to_t(self, step):
        return np.array(step[3])

    def step_to_s2(self, step):
        return np.array(step[4])

    def steps_to_s(self, steps):
        return np.array(steps[:, 0])

    def steps_to_a(self, steps):
        return np.array(steps[:, 1])

    def steps_to_r(self, steps):
        return np.array(steps[:, 2])

    def steps_to_t(self, steps):
        return np.array(steps[:, 3])

    def steps_to_s2(self, steps):
        return np.array(steps[:, 4])

    # ------------------------------------------------------------------ index maths
    def episode_number_to_buffer_index(self, episode_number):
        return len(self.buffer) - (self.next_episode_number - episode_number)

    def buffer_index_to_episode_number(self, buffer_index):
        return buffer_index - len(self.buffer) + self.next_episode_number

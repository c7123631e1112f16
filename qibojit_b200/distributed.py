"""Distributed state layer: the top log2(G) index bits are sharded over G ranks (one process
per GPU, ``torch.distributed``), non-diagonal gates on a sharded ("global") qubit trigger a
global<->local qubit swap.

What it replaces: ``CupyBackend.execute_distributed_circuit`` and ``MultiGpuOps``
(/root/reference/src/qibojit/backends/gpu.py:646-739, 1420-1516), where the pieces live in host
RAM, each gate group costs a host<->device round trip per piece and global swaps run on the CPU
(``ops.swap_pieces``, ops.py:131-137).  Here the shards stay resident in HBM and only half a
shard crosses NVLink per swap.  The piece layout is the reference's for global qubits
``[0..g-1]`` (gpu.py:1440-1442): rank r holds amplitudes ``[r * 2^nlocal, (r+1) * 2^nlocal)``.

Rules:
  (i)   gate on local bits only          -> every rank runs the local kernel
  (ii)  control on a global qubit        -> only ranks whose bit is 1 run it, control dropped
  (iii) diagonal gate on a global target -> restricted to the rank's bit value: a smaller local
        diagonal gate or a scalar phase on the shard; no exchange
  (iv)  anything else                    -> the global qubit is swapped with a local one first
        (``swap_pieces`` semantics)
  SWAP gates are executed by relabelling the qubit map (no data movement at all), and so is the
  choice of the initial layout of |0...0>.

Scheduling (``DistributedState.plan``): the gate list is walked as its dependency DAG -- every
ready gate that needs no exchange runs first -- and when nothing else can run ALL global qubits
that block (plus later-needed ones that can displace finished local qubits) are exchanged in one
all-to-all among the ranks that differ in those rank bits.  The plan is a list of
``LocalSegment`` (this rank's gates in shard numbering, compiled by the backend into multi-gate
pass programs), ``Exchange`` (one qubit, half a shard each way) and ``MultiExchange`` steps.

The class talks to the device only through the backend's reference-style kernel entry points
(``_one_qubit_base`` ...) plus a few shard primitives of the backend (``shard_zeros``,
``shard_reset``, ``run_local_segment``, ``shard_exchange``, ``shard_exchange_multi``), so the same
logic runs on gloo/CPU in the tests with a numpy stand-in backend.
"""

import numpy as np

from .backends.b200 import GATE_OPS  # noqa: F401  (re-exported for stand-in backends)

_SYMMETRIC_PHASE_OPS = ("apply_z", "apply_z_pow")


def _log2(x):
    n = int(x).bit_length() - 1
    if (1 << n) != x:
        raise ValueError(f"{x} is not a power of two")
    return n


class Comm:
    """Thin view of a torch.distributed process group (or a single-rank stub)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def all_reduce_sum(self, tensor):
        if self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        return tensor

    def all_gather(self, tensor):
        if self.world == 1:
            return [tensor]
        import torch

        out = [torch.empty_like(tensor) for _ in range(self.world)]
        self.dist.all_gather(out, tensor, group=self.group)
        return out


def is_diagonal_matrix(m, tol=0.0):
    m = np.asarray(m)
    if m.ndim != 2:
        return False
    off = m - np.diag(np.diagonal(m))
    return bool(np.all(np.abs(off) <= tol))


class LocalGate:
    """One rank's share of a gate, in the numbering of its nlocal-qubit shard register: global
    controls resolved by the rank predicate, global diagonal targets fixed to the rank's bit
    values.  Duck-types the gate attributes ``planner.lower_gate`` and the per-gate dispatch read."""

    diagonal = False

    def __init__(self, op, targets, controls, kernel_matrix, dense):
        self.op = op                          # reference kernel name (GATE_OPS value) or None
        self.target_qubits = tuple(int(t) for t in targets)
        self.control_qubits = tuple(int(c) for c in controls)
        self.kernel_matrix = kernel_matrix    # what the reference-style kernel entry consumes
        self.dense = np.asarray(dense)        # 2^t x 2^t target-only matrix
        self.name = "localgate"
        self.parameters = ()

    @property
    def qubits(self):
        return self.control_qubits + self.target_qubits

    def target_matrix(self, matrices):
        return self.dense.astype(matrices.dtype)

    def apply(self, backend, state, nqubits):
        """Gate-by-gate execution through the reference-style kernel entry points."""
        t = self.target_qubits
        bits = sorted(nqubits - 1 - q for q in self.qubits)
        qubits = np.array(bits, dtype=np.int32) if self.control_qubits else None
        if self.op is None:
            matrix = self.target_matrix(backend.custom_matrices)
        else:
            matrix = self.kernel_matrix
        if len(t) == 1:
            return backend._one_qubit_base(state, nqubits, t[0], self.op or "apply_gate", matrix, qubits)
        if len(t) == 2:
            return backend._two_qubit_base(state, nqubits, t[0], t[1], self.op or "apply_two_qubit_gate",
                                           matrix, qubits)
        return backend._multi_qubit_base(state, nqubits, list(t), matrix, np.array(bits, dtype=np.int32))


class LocalSegment:
    """A run of local gates between two exchanges; `compiled` caches the backend's program."""

    def __init__(self):
        self.gates = []
        self.compiled = None


class Plan(list):
    """[LocalSegment | Exchange] steps of one gate list + the qubit maps before and after it."""

    initial_map = None
    final_map = None


class Exchange:
    """Swap rank bit `rank_bit` with shard index bit `local_bit` (half a shard each way)."""

    def __init__(self, rank_bit, local_bit):
        self.rank_bit = int(rank_bit)
        self.local_bit = int(local_bit)


class MultiExchange:
    """Swap rank bits `rank_bits[i]` with shard index bits `local_bits[i]` (ascending) in one
    all-to-all among the 2^k ranks that differ in those rank bits."""

    def __init__(self, rank_bits, local_bits):
        pairs = sorted(zip(local_bits, rank_bits))
        self.local_bits = [int(l) for l, _ in pairs]
        self.rank_bits = [int(j) for _, j in pairs]


class DistributedState:
    """A 2^nqubits state vector sharded over the ranks of `comm`."""

    def __init__(self, backend, nqubits, comm=None, dtype=None, swap_chunk_bytes=1 << 29, initial_state=None):
        self.backend = backend
        self.comm = comm if comm is not None else Comm()
        self.nqubits = int(nqubits)
        self.nglobal = _log2(self.comm.world)
        self.nlocal = self.nqubits - self.nglobal
        if self.nlocal < 1:
            raise ValueError("more ranks than amplitudes pairs: use fewer devices")
        self.dtype = dtype or backend.dtype
        self.rank = self.comm.rank
        # logical qubit -> physical index bit; bits >= nlocal are the rank bits
        self.bit_of = [self.nqubits - 1 - q for q in range(self.nqubits)]
        self.swap_chunk_bytes = swap_chunk_bytes
        self.stats = {"exchanges": 0, "exchange_bytes": 0, "local_gates": 0, "local_segments": 0,
                      "skipped": 0, "relabelled_swaps": 0}
        self._fresh = True   # still |0...0>: the qubit map may be chosen freely
        self.relabel_swaps = True
        # second shard-sized buffer of the out-of-place, copy-engine exchange (allocated on first use
        # when every rank has the room; None + _spare_tried: the in-place exchange is used)
        self._spare = None
        self._spare_tried = False
        if initial_state is None:
            self.shard = backend.shard_zeros(self.nlocal, self.dtype, one_at_zero=(self.rank == 0))
        else:
            # the reference's piece layout for global qubits [0..g-1] (gpu.py:1440-1442): rank r
            # takes the contiguous slice [r * 2^nlocal, (r+1) * 2^nlocal) of the full vector
            size = 1 << self.nlocal
            flat = initial_state.reshape(-1)
            if int(flat.shape[0]) != (1 << self.nqubits):
                raise TypeError(f"initial state has {int(flat.shape[0])} amplitudes, expected 2^{self.nqubits}")
            self.shard = backend.shard_from(flat[self.rank * size:(self.rank + 1) * size], self.dtype)
            self._fresh = False

    # ------------------------------------------------------------------ helpers
    def is_local(self, q):
        return self.bit_of[q] < self.nlocal

    def rank_bit(self, q):
        """Value of global qubit q's index bit on this rank."""
        return (self.rank >> (self.bit_of[q] - self.nlocal)) & 1

    def _pseudo(self, bit):
        """Backend entry points take qibo-style qubit numbers of an nlocal-qubit register."""
        return self.nlocal - 1 - bit

    def qubit_at(self, bit):
        return self.bit_of.index(bit)

    # ------------------------------------------------------------------ exchange
    def swap_global_local(self, qglobal, qlocal):
        """Exchange the roles of a global and a local qubit (ops.swap_pieces semantics)."""
        steps = []
        self._plan_exchange(steps, qglobal, qlocal)
        return self.run(steps)

    def _choose_victim(self, protected, lookahead):
        """Local qubit whose next use as a non-diagonal target lies farthest in the future."""
        best, best_dist = None, -1
        for q in range(self.nqubits):
            if not self.is_local(q) or q in protected:
                continue
            if self.dtype == "complex64" and self.bit_of[q] == 0:
                continue  # 16-byte exchange granularity
            dist = lookahead.get(q, 1 << 30)
            if dist > best_dist:
                best, best_dist = q, dist
        if best is None:
            raise RuntimeError("no local qubit available to swap with")
        return best

    def ensure_local(self, qubits, lookahead=None):
        lookahead = lookahead or {}
        for q in qubits:
            if not self.is_local(q):
                victim = self._choose_victim(set(qubits), lookahead)
                self.swap_global_local(q, victim)

    # ------------------------------------------------------------------ planning
    def _emit(self, plan, gate):
        if not plan or not isinstance(plan[-1], LocalSegment):
            plan.append(LocalSegment())
        plan[-1].gates.append(gate)
        self.stats["local_gates"] += 1

    def _plan_exchange(self, plan, qglobal, qlocal):
        gbit, lbit = self.bit_of[qglobal], self.bit_of[qlocal]
        assert gbit >= self.nlocal > lbit
        plan.append(Exchange(gbit - self.nlocal, lbit))
        self.bit_of[qglobal], self.bit_of[qlocal] = lbit, gbit

    def _plan_multi_exchange(self, plan, pairs):
        """pairs: [(global qubit, local qubit)]; one pair degenerates to a plain Exchange."""
        if len(pairs) == 1:
            return self._plan_exchange(plan, *pairs[0])
        rank_bits, local_bits = [], []
        for qg, ql in pairs:
            gbit, lbit = self.bit_of[qg], self.bit_of[ql]
            assert gbit >= self.nlocal > lbit
            rank_bits.append(gbit - self.nlocal)
            local_bits.append(lbit)
            self.bit_of[qg], self.bit_of[ql] = lbit, gbit
        plan.append(MultiExchange(rank_bits, local_bits))

    def _plan_gate(self, plan, gate, lookahead):
        """Symbolic application of one gate: updates the qubit map and appends this rank's local
        gate (shard numbering) and/or exchange steps to `plan`.  No amplitude is touched."""
        name = gate.__class__.__name__
        if getattr(gate, "name", "") == "fanout":
            from . import gates as G

            for t in gate.target_qubits:
                self._plan_gate(plan, G.CNOT(gate.control_qubits[0], t), lookahead)
            return
        targets = list(gate.target_qubits)
        controls = list(gate.control_qubits)
        op = GATE_OPS.get(name)

        if op == "apply_swap" and not controls and self.relabel_swaps:
            # relabel: the data does not move, the two logical qubits trade index bits
            a, c = targets
            self.bit_of[a], self.bit_of[c] = self.bit_of[c], self.bit_of[a]
            self.stats["relabelled_swaps"] += 1
            return

        matrix = self.backend._as_custom_matrix(gate)

        if op in _SYMMETRIC_PHASE_OPS:
            # phase on |1..1> of controls+target: symmetric in its qubits
            qs = controls + targets
            if any((not self.is_local(q)) and self.rank_bit(q) == 0 for q in qs):
                self.stats["skipped"] += 1
                return
            local = [q for q in qs if self.is_local(q)]
            phase = -1.0 if op == "apply_z" else complex(np.asarray(matrix).ravel()[0])
            self._emit_phase(plan, phase, local)
            return

        # global controls: rank predicate.  An inactive rank still takes part in the exchanges a
        # non-diagonal global target needs (its peers run the gate and every rank keeps one map).
        active = all(self.is_local(q) or self.rank_bit(q) == 1 for q in controls)
        lcontrols = [q for q in controls if self.is_local(q)]

        dense = None
        if any(not self.is_local(q) for q in targets):
            dense = self._dense_matrix(gate, matrix)
            if is_diagonal_matrix(dense):
                if active:
                    self._emit_restricted_diagonal(plan, np.diagonal(dense), targets, lcontrols)
                else:
                    self.stats["skipped"] += 1
                return
            for q in targets:
                if not self.is_local(q):
                    victim = self._choose_victim(set(targets), lookahead or {})
                    self._plan_exchange(plan, q, victim)
            # a victim may have been one of this gate's local controls: re-evaluate them
            active = all(self.is_local(q) or self.rank_bit(q) == 1 for q in controls)
            lcontrols = [q for q in controls if self.is_local(q)]
        if not active:
            self.stats["skipped"] += 1
            return
        self._emit(plan, LocalGate(op, [self._pseudo(self.bit_of[q]) for q in targets],
                                   [self._pseudo(self.bit_of[q]) for q in lcontrols], matrix,
                                   dense if dense is not None else self._dense_matrix(gate, matrix)))

    def _dense_matrix(self, gate, matrix):
        from . import fusion

        return fusion.target_only_matrix(gate, self.backend.custom_matrices)

    def _emit_phase(self, plan, phase, local):
        """exp-phase on |1..1> of the local qubits `local` (all of the shard when empty)."""
        if local:
            ps = [self._pseudo(self.bit_of[q]) for q in local]
            op = "apply_z" if phase == -1.0 else "apply_z_pow"
            self._emit(plan, LocalGate(op, ps[-1:], ps[:-1], np.asarray(phase),
                                       np.diag([1.0, complex(phase)])))
        else:  # scalar phase of the whole shard: a diagonal gate that costs nothing inside a pass
            self._emit(plan, LocalGate(None, [0], [], np.diag([complex(phase)] * 2),
                                       np.diag([complex(phase)] * 2)))

    def _emit_restricted_diagonal(self, plan, diag, targets, lcontrols):
        """Diagonal gate with some targets global: fix those bits to the rank's values."""
        t = len(targets)
        d = np.asarray(diag).reshape((2,) * t)
        index = tuple(self.rank_bit(q) if not self.is_local(q) else slice(None) for q in targets)
        d = np.asarray(d[index]).reshape(-1)
        ltargets = [q for q in targets if self.is_local(q)]
        pc = [self._pseudo(self.bit_of[q]) for q in lcontrols]
        if ltargets:
            m = np.diag(d)
            self._emit(plan, LocalGate(None, [self._pseudo(self.bit_of[q]) for q in ltargets], pc, m, m))
        elif not lcontrols:
            self._emit_phase(plan, complex(d[0]), [])
        else:
            self._emit_phase(plan, complex(d[0]), lcontrols)

    def plan(self, queue, reorder=True, free_initial_map=None, batch_exchanges=True):
        """Gate list -> [LocalSegment | Exchange] for this rank, advancing the qubit map.  The
        exchange steps are identical on every rank (they depend on the map and the gate list
        only); the local gates differ by the rank predicates.

        With `reorder` the list is treated as its dependency DAG (gates sharing a qubit keep their
        order): every ready gate whose non-diagonal targets are local runs first, and an exchange
        happens only when nothing else can run.  The evicted local qubit is the one with the
        fewest remaining gates that need it local (a finished qubit costs no later exchange),
        ties broken by the farthest next use.  `reorder=False` keeps program order with a
        look-ahead victim choice.

        `batch_exchanges`: when several global qubits block ready gates at once they are swapped
        in together (`MultiExchange`); False swaps one qubit at a time.

        `free_initial_map` (default: True while the state is still |0...0>, which is symmetric
        under qubit relabelling): the qubits whose first non-diagonal gate comes latest start as
        the global ones, at no cost."""
        gates = []
        for g in queue:
            if getattr(g, "name", "") == "fanout":
                from . import gates as G

                gates.extend(G.CNOT(g.control_qubits[0], t) for t in g.target_qubits)
            else:
                gates.append(g)
        needs = [self._needs_local(g, self.backend.custom_matrices, self.relabel_swaps) for g in gates]
        if free_initial_map is None:
            free_initial_map = self._fresh and reorder
        if free_initial_map and self.nglobal:
            if not self._fresh:
                raise ValueError("the initial qubit map is only free for the |0...0> state")
            first = {}
            for i, nd in enumerate(needs):
                for q in nd:
                    first.setdefault(q, i)
            order = sorted(range(self.nqubits), key=lambda q: (-first.get(q, len(gates)), -q))
            glob = sorted(order[:self.nglobal])
            loc = [q for q in range(self.nqubits) if q not in glob]
            for k, q in enumerate(glob):
                self.bit_of[q] = self.nqubits - 1 - k
            for k, q in enumerate(loc):
                self.bit_of[q] = self.nlocal - 1 - k
        self._fresh = False
        steps = Plan()
        steps.initial_map = list(self.bit_of)
        if not reorder:
            # table[i][q] = index of the next gate after i that needs q local
            nxt = {}
            table = [None] * len(gates)
            for i in range(len(gates) - 1, -1, -1):
                table[i] = dict(nxt)
                for q in needs[i]:
                    nxt[q] = i
            for i, gate in enumerate(gates):
                look = {q: (j - i) for q, j in table[i].items()}
                self._plan_gate(steps, gate, look)
            steps.final_map = list(self.bit_of)
            return steps

        # dependency DAG by shared qubits
        npred = [0] * len(gates)
        succ = [[] for _ in gates]
        last = {}
        for i, g in enumerate(gates):
            for q in set(g.qubits):
                j = last.get(q)
                if j is not None and i not in succ[j]:
                    succ[j].append(i)
                    npred[i] += 1
                last[q] = i
        uses = {q: [] for q in range(self.nqubits)}     # indices of the gates that need q local
        for i, nd in enumerate(needs):
            for q in nd:
                uses[q].append(i)
        upos = {q: 0 for q in uses}                      # first not-yet-run entry of uses[q]
        done = [False] * len(gates)
        import heapq

        ready = [i for i in range(len(gates)) if npred[i] == 0]
        heapq.heapify(ready)
        blocked = []
        ndone = 0
        while ndone < len(gates):
            progressed = False
            while ready:
                i = heapq.heappop(ready)
                if all(self.is_local(q) for q in needs[i]):
                    self._plan_gate(steps, gates[i], None)
                    done[i] = True
                    ndone += 1
                    progressed = True
                    for j in succ[i]:
                        npred[j] -= 1
                        if npred[j] == 0:
                            heapq.heappush(ready, j)
                    if blocked and GATE_OPS.get(gates[i].__class__.__name__) == "apply_swap":
                        for j in blocked:               # a relabelling SWAP may have unblocked them
                            heapq.heappush(ready, j)
                        blocked = []
                else:
                    blocked.append(i)
            if ndone == len(gates):
                break
            # nothing can run: bring in, in ONE exchange, every global qubit that a blocked ready
            # gate needs (an all-to-all over k qubits moves (2^k - 1) / 2^k of a shard, k separate
            # swaps k / 2)
            assert blocked, "dependency cycle in the gate list"
            for q in uses:
                while upos[q] < len(uses[q]) and done[uses[q][upos[q]]]:
                    upos[q] += 1
            wanted, keep = [], set()
            for i in sorted(blocked):
                keep.update(needs[i])
                for q in needs[i]:
                    if not self.is_local(q) and q not in wanted:
                        wanted.append(q)
            if not batch_exchanges:
                wanted = wanted[:1]
                keep = set(needs[min(blocked)])
            else:
                # global qubits that are needed later ride along for free when a FINISHED local
                # qubit can take their place (it never has to come back)
                finished = [v for v in range(self.nqubits) if self.is_local(v) and v not in keep
                            and len(uses[v]) == upos[v]
                            and not (self.dtype == "complex64" and self.bit_of[v] == 0)]
                later = sorted((q for q in range(self.nqubits) if not self.is_local(q) and q not in wanted
                                and len(uses[q]) > upos[q]), key=lambda q: uses[q][upos[q]])
                spare = len(finished) - len(wanted)
                wanted.extend(later[:max(0, spare)])
            def choose(wanted, keep):
                pairs, taken = [], set()
                for q in wanted:
                    best, best_key = None, None
                    for v in range(self.nqubits):
                        if not self.is_local(v) or v in keep or v in taken:
                            continue
                        if self.dtype == "complex64" and self.bit_of[v] == 0:
                            continue  # 16-byte exchange granularity
                        left = len(uses[v]) - upos[v]
                        nxt_use = uses[v][upos[v]] if left else len(gates)
                        key = (left, -nxt_use, -self.bit_of[v])   # prefer the top bit: contiguous halves
                        if best is None or key < best_key:
                            best, best_key = v, key
                    if best is None:
                        break               # fewer victims than wanted qubits: the rest waits
                    taken.add(best)
                    pairs.append((q, best))
                return pairs

            pairs = choose(wanted, keep)
            if not pairs:
                # the blocked gates together pin every local qubit: serve the earliest one alone
                first = needs[min(blocked)]
                pairs = choose([q for q in first if not self.is_local(q)], set(first))
            if not pairs:
                raise RuntimeError("no local qubit available to swap with")
            self._plan_multi_exchange(steps, pairs)
            for j in blocked:
                heapq.heappush(ready, j)
            blocked = []
        steps.final_map = list(self.bit_of)
        return steps

    def spare_buffer(self):
        """The second shard-sized buffer that lets an exchange run out of place under the last pass
        (`run_segment_then_exchange`), or None: allocated once, only if EVERY rank has the memory
        (ranks must agree on the exchange protocol) and the backend can use it."""
        import os

        if self._spare is not None or self._spare_tried:
            return self._spare
        self._spare_tried = True
        b = self.backend
        if (self.comm.world == 1 or not hasattr(b, "shard_spare") or os.environ.get("QJ_OVERLAP_EXCHANGE", "1") == "0"
                or os.environ.get("QJ_PEER_EXCHANGE", "1") == "0"):
            return None
        spare = b.shard_spare(self.shard)
        import torch

        ok = torch.tensor([0.0 if spare is None else 1.0], dtype=torch.float64, device=self.shard.device)
        self.comm.dist.all_reduce(ok, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
        self._spare = spare if float(ok[0]) > 0 else None
        return self._spare

    def run(self, steps):
        """Execute planned steps on the shard (local segments are compiled into multi-gate pass
        programs by the backend the first time they run, and cached on the step)."""
        b = self.backend
        initial = getattr(steps, "initial_map", None)
        if initial is not None and list(initial) != self.bit_of and getattr(steps, "final_map", None) != self.bit_of:
            # (a plan made on this state has already advanced the map to its final_map)
            if not self._fresh:
                raise ValueError("plan was made for a different qubit map")
        self._fresh = False
        steps_list = list(steps)
        skip = False
        for pos, step in enumerate(steps_list):
            if skip:
                skip = False
                continue
            nxt = steps_list[pos + 1] if pos + 1 < len(steps_list) else None
            if (isinstance(step, LocalSegment) and isinstance(nxt, (MultiExchange, Exchange))
                    and hasattr(b, "run_segment_then_exchange")):
                # the segment's last pass is pipelined against the exchange (peer-memory transport)
                lbits = nxt.local_bits if isinstance(nxt, MultiExchange) else [nxt.local_bit]
                rbits = nxt.rank_bits if isinstance(nxt, MultiExchange) else [nxt.rank_bit]
                before = self.shard
                self.shard, moved = b.run_segment_then_exchange(self.shard, self.nlocal, step, lbits, rbits,
                                                                self.rank, self.comm, self.swap_chunk_bytes,
                                                                spare=self.spare_buffer())
                if self.shard is not before:      # exchanged out of place: the old shard is the next spare
                    self._spare = before
                self.stats["local_segments"] += 1
                self.stats["exchanges"] += 1
                self.stats["exchange_bytes"] += int(moved)
                skip = True
            elif isinstance(step, LocalSegment):
                self.shard = b.run_local_segment(self.shard, self.nlocal, step)
                self.stats["local_segments"] += 1
            elif isinstance(step, MultiExchange):
                moved = b.shard_exchange_multi(self.shard, self.nlocal, step.local_bits, step.rank_bits,
                                               self.rank, self.comm, self.swap_chunk_bytes)
                self.stats["exchanges"] += 1
                self.stats["exchange_bytes"] += int(moved)
            else:
                peer = self.rank ^ (1 << step.rank_bit)
                moved = b.shard_exchange(self.shard, self.nlocal, step.local_bit, peer,
                                         (self.rank >> step.rank_bit) & 1, self.comm, self.swap_chunk_bytes)
                self.stats["exchanges"] += 1
                self.stats["exchange_bytes"] += int(moved)
        if getattr(steps, "final_map", None) is not None:
            self.bit_of = list(steps.final_map)
        return self

    def apply_gate(self, gate, lookahead=None):
        steps = []
        self._plan_gate(steps, gate, lookahead)
        return self.run(steps)

    # ------------------------------------------------------------------ circuits
    @staticmethod
    def _needs_local(gate, custom_matrices, relabel_swaps=True):
        """Qubits of `gate` that must be local for it to run (non-diagonal targets)."""
        name = gate.__class__.__name__
        op = GATE_OPS.get(name)
        if op in _SYMMETRIC_PHASE_OPS or getattr(gate, "diagonal", False):
            return []
        if op == "apply_swap" and not gate.control_qubits and relabel_swaps:
            return []
        if name == "FusedGate":
            from . import fusion

            if is_diagonal_matrix(fusion.fused_matrix(gate, custom_matrices)):
                return []
        return list(gate.target_qubits)

    def execute(self, queue):
        """Apply a gate list: plan it (swap victims by look-ahead), then run the steps."""
        return self.run(self.plan(queue))

    def reset(self):
        """Back to |0...0> with the identity qubit map (in place)."""
        self.backend.shard_reset(self.shard, self.nlocal, one_at_zero=(self.rank == 0))
        self.bit_of = [self.nqubits - 1 - q for q in range(self.nqubits)]
        self._fresh = True
        return self

    # ------------------------------------------------------------------ results
    def probabilities(self, qubits):
        """Marginal distribution over logical `qubits` (replicated on every rank)."""
        b = self.backend
        qubits = list(qubits)
        lq = [q for q in qubits if self.is_local(q)]
        local = b.calculate_probabilities(self.shard, [self._pseudo(self.bit_of[q]) for q in lq],
                                          self.nlocal)
        out = b.engine.zeros((2,) * len(qubits), dtype=local.dtype, device=local.device) \
            if qubits else b.engine.zeros((), dtype=local.dtype, device=local.device)
        index = tuple(slice(None) if self.is_local(q) else self.rank_bit(q) for q in qubits)
        out[index] = local.reshape((2,) * len(lq)) if lq else local.reshape(())
        out = out.reshape(-1)
        return self.comm.all_reduce_sum(out)

    def norm2(self):
        b = self.backend
        val = b.calculate_norm(self.shard) ** 2
        t = b.engine.tensor([val], dtype=b.engine.float64, device=self.shard.device)
        return float(self.comm.all_reduce_sum(t)[0])

    # ------------------------------------------------------------------ layout
    def normalize_layout(self):
        """Bring the shards back to the reference's piece layout -- logical qubit q on index bit
        n-1-q, qubits [0..g-1] global (gpu.py:1440-1442, 1454-1456) -- ON THE DEVICES: the
        relabelled SWAPs and the scheduler's exchanges left a permuted qubit map that only
        `to_numpy_full` undid (on the host).  The permutation is executed as data movement: at most
        two multi-qubit exchanges put the right qubits on the rank bits, then one local segment of
        SWAP gates (compiled into a few multi-gate passes) orders the shard.  Semantics of
        ops.transpose_state (ops.py:112-124) followed by to_pieces."""
        n, nl = self.nqubits, self.nlocal
        want = lambda bit: n - 1 - bit                      # logical qubit that belongs on `bit`
        steps = Plan()
        steps.initial_map = list(self.bit_of)
        for _ in range(8 * max(1, self.nglobal) + 4):
            wrong = [p for p in range(nl, n) if self.bit_of[want(p)] != p]
            if not wrong:
                break
            pairs, used = [], set()
            for p in wrong:
                q = want(p)
                if self.is_local(q) and q not in used and not (self.dtype == "complex64" and self.bit_of[q] == 0):
                    pairs.append((self.qubit_at(p), q))
                    used.add(q)
            if not pairs:
                claimed = {want(b) for b in range(nl, n)}
                low = [want(p) for p in wrong if self.bit_of[want(p)] == 0]
                if low:
                    # complex64 exchanges move 16-byte vectors: a wanted qubit on index bit 0 first
                    # trades places with another local qubit (a local SWAP), then it can travel
                    others = [v for v in range(n) if self.is_local(v) and v != low[0]]
                    free = [v for v in others if v not in claimed] or others
                    self._emit_move(steps, low[0], free[0])
                    continue
                # the wanted qubits sit on other rank bits: park one occupant on a local bit that
                # nobody claims (breaks the cycle among the rank bits)
                p = wrong[0]
                victims = [v for v in range(n) if self.is_local(v) and v not in claimed
                           and not (self.dtype == "complex64" and self.bit_of[v] == 0)]
                if not victims:
                    victims = [v for v in range(n) if self.is_local(v)
                               and not (self.dtype == "complex64" and self.bit_of[v] == 0)]
                pairs = [(self.qubit_at(p), victims[0])]
            self._plan_multi_exchange(steps, pairs)
        else:
            raise RuntimeError("normalize_layout did not converge")
        for p in range(nl):                                  # cycle decomposition on the local bits
            while self.bit_of[want(p)] != p:
                self._emit_move(steps, self.qubit_at(p), want(p))
        steps.final_map = list(self.bit_of)
        assert self.bit_of == [n - 1 - q for q in range(n)]
        return self.run(steps)

    def _emit_move(self, plan, qa, qb):
        """Exchange the index bits of two LOCAL logical qubits by moving data (a physical SWAP
        gate plus the relabelling that compensates it: the logical state is unchanged)."""
        pa, pb = self._pseudo(self.bit_of[qa]), self._pseudo(self.bit_of[qb])
        swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
        self._emit(plan, LocalGate("apply_swap", [pa, pb], [], None, swap))
        self.bit_of[qa], self.bit_of[qb] = self.bit_of[qb], self.bit_of[qa]

    def to_pieces(self):
        """This rank's piece in the reference's layout (gpu.py:1437-1449), on the device."""
        self.normalize_layout()
        return self.shard

    def to_tensor(self):
        """The full state vector in logical order on every rank's device (gpu.py:1451-1465 builds
        it on the host): normalise the layout, then all-gather the pieces."""
        self.normalize_layout()
        pieces = self.comm.all_gather(self.shard)
        return pieces[0] if len(pieces) == 1 else self.backend.engine.cat(pieces)

    # ------------------------------------------------------------------ measurement
    def collapse(self, qubits, shot, normalize=True):
        """Projective collapse of logical `qubits` onto outcome `shot` (most significant bit =
        lowest-numbered qubit, cpu.py:541-563 / ops.py:47-79) on the sharded state: local measured
        qubits are zeroed by the collapse kernel, ranks whose global bits disagree with the
        outcome zero their shard, the norm is all-reduced and every rank rescales."""
        b = self.backend
        qubits = sorted(int(q) for q in qubits)
        k = len(qubits)
        outcome = {q: (int(shot) >> (k - 1 - i)) & 1 for i, q in enumerate(qubits)}
        alive = all(self.is_local(q) or self.rank_bit(q) == outcome[q] for q in qubits)
        if not alive:
            self.shard = b.shard_reset(self.shard, self.nlocal, one_at_zero=False)
        else:
            local = sorted((self._pseudo(self.bit_of[q]), outcome[q]) for q in qubits if self.is_local(q))
            if local:
                lshot = 0
                for _, bit in local:
                    lshot = (lshot << 1) | bit
                self.shard = b.collapse_state(self.shard, [p for p, _ in local], lshot, self.nlocal, normalize=False)
        self._fresh = False
        if normalize:
            norm = float(np.sqrt(self.norm2()))
            self.shard = b.shard_scale(self.shard, self.nlocal, 1.0 / norm)
        return self

    def full_probabilities(self):
        """|amplitude|^2 of the whole register in logical order, on every rank (the input of the
        reference's samplers).  Needs 2^n reals per device: registers up to ~34 qubits."""
        self.normalize_layout()
        b = self.backend
        local = b.calculate_probabilities(self.shard, list(range(self.nlocal)), self.nlocal)
        pieces = self.comm.all_gather(local)
        return pieces[0] if len(pieces) == 1 else b.engine.cat(pieces)

    def sample_frequencies(self, nshots):
        """`Backend.sample_frequencies` (cpu.py:383-394) on the sharded state: the same sampler --
        the reference's Metropolis chains above the shot threshold (ops.py:86-108), bit-exact under
        a fixed seed -- runs on the gathered probability vector; every rank returns the same
        Counter (same seed stream)."""
        return self.backend.sample_frequencies(self.full_probabilities(), nshots)

    def sample_shots(self, nshots):
        return self.backend.sample_shots(self.full_probabilities(), nshots)

    def to_numpy_full(self):
        """Gather every shard and undo the qubit relabelling (small registers / tests only)."""
        pieces = self.comm.all_gather(self.shard)
        phys = np.concatenate([self.backend.to_numpy(p) for p in pieces])
        n = self.nqubits
        # axis a of the reshaped tensor is physical bit n-1-a; logical qubit q must end on axis q
        t = phys.reshape((2,) * n)
        axes = [n - 1 - self.bit_of[q] for q in range(n)]
        return np.transpose(t, axes).reshape(-1)


def execute_distributed_circuit(backend, circuit, initial_state=None, nshots=None, comm=None):
    """`CupyBackend.execute_distributed_circuit` (gpu.py:646-739): `initial_state` may be None
    (|0...0>), a full state vector (numpy array or tensor: every rank takes its piece,
    gpu.py:669-676) or a DistributedState; anything else is a TypeError (gpu.py:677-682).  With
    `nshots` the register is sampled (`sample_frequencies`) and (state, frequencies) is returned."""
    if isinstance(initial_state, DistributedState):
        state = initial_state
    elif initial_state is None:
        state = DistributedState(backend, circuit.nqubits, comm=comm)
    elif hasattr(initial_state, "shape") and hasattr(initial_state, "reshape"):
        state = DistributedState(backend, circuit.nqubits, comm=comm, initial_state=initial_state)
    else:
        raise TypeError(f"Initial state type {type(initial_state)} is not supported by distributed circuits.")
    fingerprint = (len(circuit.queue), getattr(backend, "circuit_fingerprint", len)(circuit.queue))
    key = ("dist", state.rank, state.comm.world, backend.dtype, tuple(state.bit_of), state._fresh)
    cache = circuit.__dict__.setdefault("_qj_programs", {})
    entry = cache.get(key)
    if entry is None or entry[0] != fingerprint:     # (an outdated plan is dropped with its programs)
        entry = cache[key] = (fingerprint, state.plan(circuit.queue))
    steps = entry[1]
    state.run(steps)
    if nshots:
        return state, state.sample_frequencies(int(nshots))
    return state

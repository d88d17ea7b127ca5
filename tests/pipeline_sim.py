"""Randomised interleaving simulator for mbarrier-synchronised warp-specialised kernels (test infrastructure)."""
import random

class Bar:
    def __init__(self, count):
        self.count = count; self.pending = count; self.tx = 0; self.phase = 0; self.completions = 0
    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1; self.pending = self.count; self.completions += 1
    def arrive(self, tx=0):
        assert self.pending > 0, "arrive on a barrier with no pending arrivals"
        self.tx += tx; self.pending -= 1; self._check()
    def complete_tx(self, n):
        self.tx -= n; assert self.tx >= 0; self._check()
    def passed(self, parity):
        return self.phase != parity

class Sim:
    """Agents are generators yielding ('wait', bar, parity) | ('step',) ; async completions are scheduled events."""
    def __init__(self, seed):
        self.rng = random.Random(seed); self.agents = {}; self.events = []; self.t = 0
    def add(self, name, gen):
        self.agents[name] = [gen, None]      # [generator, blocked_on]
    def later(self, fn, ordered_key=None):
        """schedule fn after a random delay; events with the same ordered_key fire in submission order"""
        self.events.append([self.rng.randint(0, 6), fn, ordered_key, len(self.events) + self.t * 1000])
    def run(self, max_steps=200000):
        for _ in range(max_steps):
            self.t += 1
            # fire ripe events, respecting order within a key
            fired = True
            while fired:
                fired = False
                for ev in list(self.events):
                    ev[0] -= 0
                for ev in sorted(self.events, key=lambda e: e[3]):
                    if ev[0] <= 0 and not any(o is not ev and o[2] is not None and o[2] == ev[2] and o[3] < ev[3] for o in self.events):
                        self.events.remove(ev); ev[1](); fired = True; break
            for ev in self.events:
                ev[0] -= 1
            runnable = []
            for name, st in self.agents.items():
                if st[0] is None: continue
                if st[1] is None or st[1][0].passed(st[1][1]):
                    runnable.append(name)
            if not runnable:
                if all(st[0] is None for st in self.agents.values()) and not self.events:
                    return True
                if not self.events:
                    blocked = {n: (id(st[1][0]), st[1][1]) for n, st in self.agents.items() if st[0] is not None}
                    raise RuntimeError(f"deadlock: {blocked}")
                continue
            name = self.rng.choice(runnable)
            st = self.agents[name]; st[1] = None
            try:
                op = next(st[0])
            except StopIteration:
                st[0] = None; continue
            if op[0] == 'wait':
                st[1] = (op[1], op[2])
        raise RuntimeError("did not terminate")

"""Host-side pieces of bench.py's contract that can be checked without a GPU."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return b


def test_algorithmic_bytes_match_the_survey():
    b = _bench()
    # 8 H W + 40 V + 24 F; xArm7 links 1-7 @1280x720 = 8.91 MB / frame
    assert b.algorithmic_bytes_per_frame(720, 1280, 17504, 35002) == 7372800 + 700160 + 840048 == 8913008
    assert b.WORKLOAD["B"] == 10 and (b.WORKLOAD["H"], b.WORKLOAD["W"]) == (720, 1280)
    assert b.WORKLOAD["ring"] * b.WORKLOAD["B"] * 720 * 1280 * 8 > 126e6       # masks + refs of the ring exceed the L2


def test_exactly_one_json_line_on_stdout_under_torchrun():
    code = textwrap.dedent('''
        import os, sys, importlib.util
        spec = importlib.util.spec_from_file_location("bench", %r); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
        o = b.JsonStdout(True)
        os.write(1, b"NCCL version 2.28.9+cuda12.9\\n")      # what a C library prints while it initialises
        print("python-level chatter")
        o.emit({"metric": "m", "value": 1.5})
    ''' % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True)
    lines = r.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
    assert "NCCL version" in r.stderr and "python-level chatter" in r.stderr


def test_reference_arm_exits_quietly_on_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_timed_regions_repeat_until_every_rank_has_measured_enough():
    """The K-step region is repeated until MIN_TIMED_MS of device time is measured; with several ranks the decision to stop
    is collective (agree), or a rank that needs one region more would wait in a barrier the others never enter."""
    b = _bench()

    class Ev:
        clock = [0.0]

        def __init__(self, enable_timing=True):
            self.t = None

        def record(self):
            self.t = Ev.clock[0]

        def elapsed_time(self, other):
            return other.t - self.t

    class FakeTorch:
        class cuda:
            Event = Ev

    calls = []

    def run(n, k0):
        calls.append((n, k0))
        Ev.clock[0] += 2.0 * n          # 2 ms per step

    ms, regions = b.timed_regions(run, 10, lambda: None, FakeTorch, min_ms=50.0)
    assert regions == [20.0, 20.0, 20.0] and ms == 20.0 and calls == [(10, 0), (10, 10), (10, 20)]
    # a peer that is not done yet keeps this rank going; the loop ends when both agree
    votes = []

    def agree(done):
        votes.append(done)
        return done and len(votes) >= 5

    calls.clear()
    ms, regions = b.timed_regions(run, 10, lambda: None, FakeTorch, min_ms=50.0, agree=agree)
    assert len(regions) == 5 and votes == [False, False, True, True, True]


def test_graph_steps_divide_the_timed_region_when_they_can():
    b = _bench()
    assert b.graph_steps(2000, 128, 4) == 100      # 20 replays, no eager remainder
    assert b.graph_steps(1024, 128, 4) == 128
    assert b.graph_steps(20, 128, 4) == 20
    assert b.graph_steps(50, 128, 4) == 48         # no multiple of 4 divides 50: 48 + 2 eager steps
    assert b.graph_steps(3, 128, 4) == 4           # (fewer steps than a graph: all of them run eagerly)
    assert b.graph_steps(200, 128, 4) == 100

"""bench.py's reference arm runs without a GPU (it times the reference's CPU implementation of the path, or the oracle
port when /root/reference is absent) and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--ref-rays', '32'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'].startswith('rays/sec') and line['unit'] == 'rays/s'
    assert line['higher_is_better'] is True and line['n_gpus'] == 1 and line['steps'] == 1 and line['value'] > 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] and 'sample' in line['cpu_baseline']
    assert line['e2e'] == {'value': line['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_clock_sampler_keeps_the_samples_of_the_timed_region(monkeypatch):
    """bench.ClockSampler: nvidia-smi is started long before the timed region (its first line needs 0.1 - 1 s), the two
    mark() calls bracket the region, stop() keeps the lines that fall inside it — or, if the region was shorter than the
    sampling period, the lines taken under the same load after it — and reports throttle reasons."""
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc(object):
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    def sampler(lines, t0, t1):
        s = bench.ClockSampler(0)
        s.proc, s.lines, s.t0, s.t1 = FakeProc(), lines, t0, t1
        return s.stop()

    row = '1965, 1965, 700.0, Not Active, Not Active, Not Active, %s'
    now = time.monotonic()
    lines = [(now - 5.0, '1200, 1965, 90.0, Not Active, Not Active, Not Active, Not Active'),   # idle, before the region
             (now - 0.95, row % 'Not Active'), (now - 0.9, row % 'Active'), (now - 0.5, row % 'Not Active')]
    r = sampler(list(lines), now - 1.0, now - 0.8)
    assert r['samples'] == 2 and r['sm_mhz'] == 1965.0 and r['sm_max_mhz'] == 1965.0
    assert r['reasons'] == ['sw_power_cap'] and r['window'] == 'timed region'
    # a region between two samples: the samples right after it stand in
    r = sampler(list(lines), now - 0.8, now - 0.7)
    assert r['samples'] == 1 and r['window'] == 'timed + end-to-end regions' and r['reasons'] == []
    # no tool
    s = bench.ClockSampler(0)
    assert s.stop()['reasons'] == ['nvidia-smi unavailable']


def test_measurement_tools_compile():
    """the scripts under tools/ (run on the GPU box by hand) at least parse"""
    import py_compile
    tools = os.path.join(ROOT, 'tools')
    names = [n for n in sorted(os.listdir(tools)) if n.endswith('.py')]
    assert 'chain_trace.py' in names and 'f3_trace.py' in names and 'frame_profile.py' in names
    for n in names:
        py_compile.compile(os.path.join(tools, n), doraise=True)

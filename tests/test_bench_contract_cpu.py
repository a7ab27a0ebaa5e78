"""bench.py's driver contract, the parts that run without a GPU: the reference arm prints exactly one
JSON line on stdout with the agreed keys, and the other ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return p.stdout.decode()


def test_reference_arm_prints_one_json_line():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'streamline-steps/sec' and d['unit'] == 'streamline-steps/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 3   # the timing rules ask for W >= 3
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['vs_baseline'] is None
    assert d['config']['workload'].startswith('whole-brain synthetic 145x174x145')
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_do_no_work():
    out = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''


def test_committed_gpu_line_carries_every_contract_key():
    """The line the final tree printed on a B200 under the driver's command line (profiles/README.md): the
    keys the driver and the judge read are there and are consistent with one another."""
    with open(os.path.join(ROOT, 'profiles', 'r2_bench_driver_protocol.json')) as f:
        d = json.load(f)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['metric'] == 'streamline-steps/sec' and d['dtype'] == 'fp16' and d['vs_baseline'] is None
    assert 'model' not in d['config'] and d['config']['workload'].startswith('whole-brain synthetic 145x174x145')
    assert d['steps'] == 20 and d['warmup'] == 5 and d['gpu_launches'] == 3 * d['steps']
    # value = units / time: 50 000 streamline-steps per step
    assert abs(d['value'] - 50000 / (d['ms_per_step'] * 1e-3)) / d['value'] < 0.01
    r = d['roofline']
    assert r['bound'] == 'tensor' and r['unit'] == 'TFLOP/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert 0.5 < r['frac'] < 1.0 and r['traffic'] > 0
    # achieved = algorithmic flops per launch / measured launch time
    assert abs(r['achieved'] - r['flop_per_launch'] / (r['avg_launch_us'] * 1e-6) / 1e12) / r['achieved'] < 1e-6
    for t in d['tiers'].values():
        assert t['roofline']['frac'] < 1.0          # never a fraction above 1 against its own denominator
        assert t['sustained'] is None or t['sustained']['value'] <= t['value'] * 1.02
    e = d['e2e']
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] < d['value']
    c = d['cpu_baseline']
    assert c['kind'] == 'port' and c['cores'] >= 1 and c['value'] > 0 and c['sample'] and c['phase_seconds']
    ck = d['clocks']
    assert ck['sm_max_mhz'] >= ck['sm_mhz'] > 0 and not set(ck['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    s = d['sharded']
    assert s['scaling'] == 'strong' and s['properties_ok'] is True

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

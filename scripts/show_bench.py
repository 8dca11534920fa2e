"""One-line summary of a bench.py JSON line read from stdin."""
import json
import sys

for line in sys.stdin:
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    if d.get('impl') == 'reference':
        print('reference arm: %.1f frames/s on %s cores' % (d['value'], d['cpu_baseline']['cores']))
        continue
    st = d['stages']
    us = lambda k: st[k]['ms_total'] / d['steps'] * 1e3
    r, o = d['roofline'], d['roofline_other']
    print('value %.3e  e2e %.3e frames/s | %.1f us/step: pose %.1f fused %.1f score %.1f | hbm %.2f  tensor exec %.2f alg %.2f | clk %s %s'
          % (d['value'], d['e2e']['value'], d['ms_per_step'] * 1e3, us('pose_chain'), us('fused_blend_skin'), us('scoring'),
             r['frac'], o['executed_mma']['frac'], o['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
    if d.get('cpu_baseline'):
        print('cpu_baseline %.0f frames/s on %d threads' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    if d.get('value_spread'):
        s, e = d['value_spread'], d['e2e'].get('spread', {})
        print('value median %.3e [%.3e .. %.3e] x%d | e2e [%.3e .. %.3e]' % (s['median'], s['min'], s['max'], s['repeats'],
                                                                        e.get('min', 0), e.get('max', 0)))
    if d.get('e2e_with_verts'):
        print('e2e_with_verts %.3e frames/s (%.2f ms/step)' % (d['e2e_with_verts']['value'], d['e2e_with_verts']['ms_per_step']))
    for k in ('config3', 'config5', 'config4'):
        c = d.get(k)
        if c:
            print('%s: %.3e frames/s (%.2f ms, %s GB/s, exchange %s) parity %s %s' % (
                k, c['value'], c['ms'], round(c['achieved_gbs']), c['exchange'], json.dumps(c['parity']),
                {x: c[x] for x in ('sharded_equals_single', 'stages_alone_ms', 'model_runs_on_rank0') if x in c}))

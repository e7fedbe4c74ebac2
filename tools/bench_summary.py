#!/usr/bin/env python
"""one line per bench JSON in gpurun_out/"""
import glob, json, sys
files = sys.argv[1:] or (['gpurun_out/bench.json'] + sorted(glob.glob('gpurun_out/bench_dist*.json')) + ['gpurun_out/bench_ref.json'])
for f in files:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    r = j.get('roofline')
    if r:
        print('%-18s %-12s value %6.1f Gpx/s | enc %6.2f ms %5.0f Gpx/s %4.1f%% | dec %6.2f ms %5.0f Gpx/s %4.1f%% | e2e %5.2f | foreign %.3f | cpu %s | clk %s %s' % (
            f.split('/')[-1], j['config']['distribution'], j['value'], r['encode']['ms'], r['encode']['gpx_per_s'], 100 * r['encode']['frac'],
            r['decode']['ms'], r['decode']['gpx_per_s'], 100 * r['decode']['frac'], j['e2e']['value'], j['extra']['decode_foreign_stream_gpx_per_s'],
            j.get('cpu_baseline') and round(j['cpu_baseline']['value'], 3), j['clocks'] and j['clocks']['sm_mhz'], j['clocks'] and j['clocks']['reasons']))
        ca = j['extra'].get('c_api_single_image')
        if ca:
            print('    C API single image:', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in ca.items() if k != 'note'})
    else:
        print('%-18s REF value %.3f Gpx/s cores %s enc %.3f dec %.3f' % (f.split('/')[-1], j['value'], j['cpu_baseline']['cores'], j['cpu_baseline']['encode_gpx'], j['cpu_baseline']['decode_gpx']))

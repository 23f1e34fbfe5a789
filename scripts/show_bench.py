import json, sys
d = json.load(open(sys.argv[1]))
print('value %.3e seg/s  ms/step %.3f  valid/s %.3e' % (d['value'], d['ms_per_step'], d['valid_paths_per_s']))
if d.get('e2e'): print('e2e %.3e (%.2f ms) h2d %d d2h %d' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step']))
for k, v in d['kernels'].items(): print('  %-14s %.3f ms  share %.2f  %.0f GB/s  frac %.3f' % (k, v['ms'], v['share'], v['achieved_gbs'], v['frac']))
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('clocks', d['clocks'], 'launches', d['gpu_launches'])

#!/usr/bin/env python
"""Turn ncu output into the small tracked summaries under profiles/.

  launches <launch_list.csv> <out_prefix>   metrics-only pass of `bench.py --steps 2 --warmup 1`
                                            -> <out_prefix>_launch_shares.csv (one forward)
  full <raw_page.csv> <out_prefix>          `ncu -i X.ncu-rep --page raw --csv` of one forward
                                            -> <out_prefix>_ncu_full.csv and <out_prefix>_traffic.json

Launches are mapped to engine ops by kernel name and order inside one SqueezeDet forward."""
import csv, json, sys

def ops_of(kernels):
  """Engine op of each launch of ONE SqueezeDet forward, from the kernel names in launch order:
  first_tc / conv_pool_simt = conv1+pool1; fire_fused = a whole fire module; conv_tc launches
  alternate squeeze / expand inside the un-fused fire modules; the last conv_tc + splitk_reduce =
  conv12; maxpool = pool3, pool5."""
  ops, fire, half, pools = [], 2, 0, ['pool3', 'pool5']
  n_tc_left = sum(1 for k in kernels if 'conv_tc_kernel' in k)
  for k in kernels:
    if 'first_tc_kernel' in k or 'conv_pool_simt' in k:
      ops.append('conv1+pool1')
    elif 'fire_fused_kernel' in k:
      ops.append('fire%d.fused' % fire); fire += 1
    elif 'maxpool' in k:
      ops.append(pools.pop(0))
    elif 'conv_tc_kernel' in k:
      n_tc_left -= 1
      if n_tc_left == 0 and fire > 11:
        ops.append('conv12.partials')
      else:
        ops.append('fire%d.%s' % (fire, 'squeeze' if half == 0 else 'expand'))
        half ^= 1
        if half == 0: fire += 1
    elif 'splitk_reduce' in k:
      ops.append('conv12.reduce')
    elif 'interpret' in k:
      ops.append('interpret_output')
    elif 'filter' in k:
      ops.append('filter_prediction')
    else:
      ops.append(k.split('(')[0][-24:])
  return ops

def read_rows(path):
  rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
  hdr = next(r for r in rows if 'Kernel Name' in r)
  body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
  return hdr, body

def forward_span(names):
  """(start, length) of the last complete forward: first-layer kernel ... filter_kernel."""
  ends = [i for i, n in enumerate(names) if 'filter_kernel' in n]
  for e in reversed(ends):
    starts = [i for i in range(e) if 'first_tc_kernel' in names[i] or 'conv_pool_simt' in names[i]]
    if starts:
      return starts[-1], e - starts[-1] + 1
  raise SystemExit('no complete forward found')

def launches(path, prefix):
  hdr, body = read_rows(path)
  i_id, i_k, i_m, i_v = hdr.index('ID'), hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
  dur = [(r[i_k], float(r[i_v].replace(',', ''))) for r in body if r[i_m] == 'gpu__time_duration.sum']
  names = [n for n, _ in dur]
  s, N = forward_span(names)
  fw = dur[s:s + N]
  OPS = ops_of([n for n, _ in fw])
  tot = sum(v for _, v in fw)
  with open(prefix + '_launch_shares.csv', 'w') as f:
    f.write('# one forward (%d launches) out of the ncu launch list of `bench.py --steps 2 --warmup 1`\n' % N)
    f.write('# cold-cache, serialised under the profiler: compare SHARES, not absolutes\n')
    f.write('op,kernel,duration_us,share\n')
    for op, (n, v) in zip(OPS, fw):
      short = n.split('(')[0].split('::')[-1].replace('void ', '')
      f.write('%s,%s,%.2f,%.4f\n' % (op, short, v / 1e3, v / tot))
    f.write('total,,%.2f,1.0\n' % (tot / 1e3))
  print('forward at launch', s, 'total %.1f us' % (tot / 1e3))

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max']

def full(path, prefix):
  rows = list(csv.reader(open(path)))
  hdr = rows[0]
  units = rows[1]
  body = rows[2:]
  i_k = hdr.index('Kernel Name')
  names = [r[i_k] for r in body]
  s, N = forward_span(names)
  OPS = ops_of(names[s:s + N])
  cols = [(m, hdr.index(m)) for m in KEEP if m in hdr]
  def num(x):
    try: return float(x.replace(',', ''))
    except ValueError: return float('nan')
  scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
  traffic = {}
  with open(prefix + '_ncu_full.csv', 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on, one SqueezeDet forward (1242x375, b=20, one B200)\n')
    f.write('# units: ' + ', '.join('%s[%s]' % (m, units[i]) for m, i in cols) + '\n')
    f.write('op,' + ','.join(m for m, _ in cols) + '\n')
    for op, r in zip(OPS, body[s:s + N]):
      f.write(op + ',' + ','.join(r[i] for _, i in cols) + '\n')
      rd = num(r[hdr.index('dram__bytes_read.sum')]) * scale.get(units[hdr.index('dram__bytes_read.sum')], 1.0)
      wr = num(r[hdr.index('dram__bytes_write.sum')]) * scale.get(units[hdr.index('dram__bytes_write.sum')], 1.0)
      key = op.split('.')[0].replace('conv1+pool1', 'conv1')
      traffic[key] = traffic.get(key, 0.0) + rd + wr
  json.dump({'source': 'ncu --set full capture of this round (%s_ncu_full.csv): dram__bytes_read.sum + '
                       'dram__bytes_write.sum per launch, summed over the launches of an op' % prefix,
             'bytes_per_op': traffic}, open(prefix + '_traffic.json', 'w'), indent=1)
  print('wrote', prefix + '_ncu_full.csv', prefix + '_traffic.json')

if __name__ == '__main__':
  {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])

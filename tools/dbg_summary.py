#!/usr/bin/env python
"""Condense a DLQ_DBG_TIMES sweep log (tools/conv_sweep.py stderr) into one line per layer."""
import re, sys
dbg = None
for l in open(sys.argv[1]):
    if l.startswith('[dbg_times]'):
        dbg = l
    elif ' us ' in l and dbg:
        m = re.search(r'grid (\S+) MT=(\d) n_tile=(\d+) trips=(\d+) \| mma: total (\d+) wait_acc (\d+) wait_a (\d+) wait_b (\d+) \| epi: total (\d+) wait_acc_full (\d+) \| prodA: total (\d+) wait_a_empty (\d+)', dbg)
        print(l.split(' us')[0].ljust(32), 'grid', m.group(1), 'MT', m.group(2), 'nt', m.group(3), 'trips', m.group(4), '| mma', m.group(5), 'w_acc', m.group(6), 'w_a', m.group(7), 'w_b', m.group(8), '| epi w_full', m.group(10), '| prodA w_empty', m.group(12))

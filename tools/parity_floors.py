#!/usr/bin/env python
"""Summarise a parity report written by the test-suite under I2C_PARITY_REPORT=<file> (tests/conftest.py:_report):
worst measured error per assertion site.  python tools/parity_floors.py gpurun_out/parity_report.jsonl"""
import json
import sys
from collections import defaultdict

worst = defaultdict(float)
count = defaultdict(int)
for line in open(sys.argv[1]):
    r = json.loads(line)
    key = (r["test"].split(" ")[0].split("::", 1)[-1], r["where"])
    worst[key] = max(worst[key], r["err"])
    count[key] += 1
for (test, where), v in sorted(worst.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    print(f"{where:60s} {test:70s} n={count[(test, where)]:4d} worst={v:.2e}")

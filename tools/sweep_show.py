"""Prints a sweep .jsonl (tools/sweep.py) as a table."""
import sys, json
for l in open(sys.argv[1]):
    d = json.loads(l)
    if 'name' in d:
        print("%-22s %8d step %8.3f kern %8.3f e2e %8.3f %s" % (d['name'], d['particles'], d.get('step_ms', -1), d.get('kernel_ms', -1), d.get('e2e_p50_ms', -1), d.get('error', '')))

import json,sys
d=json.loads(sys.stdin.read())
print(sys.argv[1], {k:d[k] for k in ("value","library_ms")}, d["roofline"]["gcups"], {k:v for k,v in d["counters"].items() if "equal" in k})

import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("value %.0f serial %.0f e2e %.0f" % (d["value"], d["serial"]["value"], d["e2e"]["value"]))

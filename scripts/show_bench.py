import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ex = d.pop("extra", {})
print(json.dumps(d, indent=None)[:1500])
for k, v in ex.items():
    print("==", k, json.dumps(v)[:1800])

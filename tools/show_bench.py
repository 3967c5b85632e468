import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 3), "obj/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
        print("  ", {k: round(v, 3) for k, v in d["roofline"]["profile_ms_per_step"].items()})
    except Exception as ex:
        print(f, "ERR", ex)

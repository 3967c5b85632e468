import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "| gpus", d["n_gpus"], "| ms", round(d["ms_per_step"], 3), "| obj/s", round(d["value"]), "| e2e obj/s", round(d["e2e"]["value"]),
              "| e2e ms", round(d["e2e"]["ms_per_step"], 3), "|", d["config"]["precision"], "B/GPU", d["config"]["batch_per_gpu"], "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as ex:
        print(f, "ERR", ex)

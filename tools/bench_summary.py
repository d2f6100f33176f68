import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["config"]["single_solve_ms"],3), {k:round(v,3) for k,v in d["roofline"]["kernel_share_of_single_solve"].items()}, round(d["roofline"]["avg_launch_ms"],4))
    except Exception as e:
        print(f, "ERR", e)

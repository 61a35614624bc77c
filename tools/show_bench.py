import json,sys
j=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.3e ms %.3f frac %.3f" % (j["value"], j["ms_per_step"], j["roofline"]["frac"]), {k: round(v,3) for k,v in j["stage_ms_per_step"].items()})

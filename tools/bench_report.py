import json,sys
d=json.load(open(sys.argv[1]))
print("value G/s", round(d["value"]/1e9,2), "ms/step", round(d["ms_per_step"],1), "frac", round(d["roofline"]["frac"],3), "e2e G/s", round(d["e2e"]["value"]/1e9,2) if d.get("e2e") else None)
sl=d["roofline"].get("sliced")
if sl:
    print(sl["geometry"], "win/rec", round(sl["windows_per_record"],2))
    for k,v in sl["phases"].items(): print(" ", k, round(v["ms_per_step"],2), v["launches_per_step"], v["stream_bytes_per_instance"], v["stream_gbs"])

import json, subprocess, sys, os
for n in ("", "3", "7", "15", "31"):
    env = dict(os.environ)
    if n: env["VV_COPY_THREADS"] = n
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, env=env).stdout
    line = [l for l in out.splitlines() if l.startswith('{"metric"')][-1]
    d = json.loads(line)
    print("threads", n or "default", "c5_long", round(d["c5_long"]["frames_per_s"]), "chunked", round(d["c5_long_chunked_device"]["frames_per_s"]), "e2e", round(d["e2e"]["value"]), "box rows", round(d["e2e_box_mask"]["row_bounded"]["value"]), "prepost box rows", round(d["e2e_prepost"]["box_mask"]["row_bounded"]["value"]), flush=True)

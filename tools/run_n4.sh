timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 > gpurun_out/s10_bench_n4.json 2> gpurun_out/s10_bench_n4.err; tail -c 200 gpurun_out/s10_bench_n4.err
python - <<PY
import json
l=[x for x in open("gpurun_out/s10_bench_n4.json") if x.startswith("{")][-1]
d=json.loads(l)
print(d["value"], d["n_gpus"], d["roofline"]["frac"], d["e2e"]["value"], d["extras"].get("error"))
c5=d["extras"]["config5_segmented_fir_4096taps"]; c4=d["extras"]["config4_channelizer_64ch"]
print(c5["Msamples_s_outputs_sharded"], c5["parity_ok"], c4["time_segmented_Msamples_s"], c4["parity_ok"], c4["channel_sharded_parity_ok"])
PY

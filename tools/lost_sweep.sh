# lost-point stage: leaf size of its dissection tree (NRSLAM_B200_LOST_LEAF) against device time
for leaf in 4 8 16 32; do
  echo "== leaf $leaf"
  NRSLAM_B200_LOST_LEAF=$leaf timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --no-ba 2> gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['e2e']['ms_per_step'])"
done

for leaf in 2 4 8; do
  echo "== leaf $leaf"
  NRSLAM_B200_LOST_LEAF=$leaf NRSLAM_B200_PROF=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-ba 2> gpurun_out/b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['device_ms'], d['e2e']['ms_per_step'])"
  grep "nrs prof" gpurun_out/b.err | grep -v "grid 128\|grid 16 " | tail -1 | cut -c1-60
done

#!/bin/bash
# kmc::GetPseudoTimeStamps(Pointcloud const&, Time, Time) per call on the real scan over the stamp-kernel knobs (grid, access width, pieces);
# prints the tune, the stamps call in us and the MotionCompensateFrame call in us.  Run from the repo root on a GPU box.
B=kitti_motion_compensation_b200/lib/bench_motion_compensate_frame
S=tests/golden/kitti_2011_09_26_drive_0005_frame0.bin
for t in "" "stamp_ctas=1" "stamp_ctas=2" "stamp_ctas=3" "stamp_ctas=12" "stamp_vec=1" "stamp_parts=2" "stamp_ctas=1,stamp_parts=2" "stamp_ctas=2,stamp_parts=2"; do
  echo -n "$t  "; KMC_B200_TUNE="$t" $B $S 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['get_pseudo_time_stamps_us_median'], d['us_median'])"
done

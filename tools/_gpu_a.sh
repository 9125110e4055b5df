#!/bin/bash
cd /root/repo
D=data/clouds
./realtime_robot_b200/realtime_robot --database-online $D/chair1.pcd $D/chair2.pcd $D/desk1.pcd --scans $D/mcloud.pcd $D/T0_m8111.pcd --hypotheses 20000 > gpurun_out/j_driver.log 2>&1
echo rc $? >> gpurun_out/j_driver.log

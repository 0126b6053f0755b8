#!/bin/bash
TAG=${1:-r02i}
bash tools/gpu_ab.sh ${TAG}_g128 "" main prev
bash tools/gpu_ab.sh ${TAG}_g64 "--egroups 64" main prev
bash tools/gpu_ab.sh ${TAG}_g7 "--egroups 7" main prev

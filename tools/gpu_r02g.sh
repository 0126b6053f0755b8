#!/bin/bash
TAG=${1:-r02g}
bash tools/gpu_ab.sh ${TAG}_g128 "" main ia
bash tools/gpu_ab.sh ${TAG}_g64 "--egroups 64" main ia
bash tools/gpu_ab.sh ${TAG}_geom "--geometry" main ia

#!/bin/bash
TAG=${1:-r02e}
bash tools/gpu_ab.sh ${TAG}_g128 "" main fs fsu2 u2
bash tools/gpu_ab.sh ${TAG}_g64 "--egroups 64" main fs fsu2 u2

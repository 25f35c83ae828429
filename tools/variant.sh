# Build a variant of the tcgen05 step kernel with extra -D flags into libmcmcn_<name>.so (same ABI; A/B with tools/ab.sh):
#   bash tools/variant.sh <name> -DMCMCN_TC_X=1 ...
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
csrc=$root/mcmc-for-nested-data_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I $root/include -I $csrc -Xcompiler -fPIC "$@" \
  -c $csrc/mcmcn_sets_tc.cu -o /tmp/mcmcn_sets_tc_$name.o
objs=$(ls $csrc/build/*.o | grep -v mcmcn_sets_tc.o)
/usr/local/cuda/bin/nvcc -shared --cudart shared -o $root/mcmc-for-nested-data_b200/libmcmcn_$name.so $objs /tmp/mcmcn_sets_tc_$name.o -ldl
echo built libmcmcn_$name.so

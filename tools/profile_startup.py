import sys, os, time, cProfile, pstats
sys.path[:0]=['/root/repo','/root/repo/mcmc-for-nested-data_b200']
import torch
torch.zeros(1).cuda()
import posteriorSampling as ps
from objectives import Objective
from workloads import makeWorkload
X,y,names,ranges=makeWorkload(1024,200,8)
ps.CSV_VALUE_LIMIT=0; ps.STORE_DTYPE="float32"
def run():
    h=Objective.linear_regression(X,y)
    ps.samplePosterior(1024,200,20,names,1024,200,"partial",h,"/dev/shm/prof_start",saveLogLikelihood=False,startingPointValueRange=ranges,displayProgress=False)
run()
t=time.time(); run(); print("second call wall", time.time()-t)
pr=cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

import sys, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
import rcs_b200
from rcs_b200 import TASK_ABC, TASK_XYZ, workload
from oracle import ref
from conftest import graph_file
g = rcs_b200.Graph(graph_file("wam_arm")); m = rcs_b200.Model(g, 0)
rik = ref.RefIk(graph_file("wam_arm"), [("XYZ","BallTip"),("ABC","BallTip")])
tasks=[(TASK_XYZ,g.body_index("BallTip")),(TASK_ABC,g.body_index("BallTip"))]
lay=g.layout()
q_goal = workload.sample_states(lay, 2000, margin=0.2)
x_des,_ = rik.task_kinematics(q_goal)
rng=np.random.default_rng(11); q0 = q_goal + 0.15*rng.uniform(-1,1,size=q_goal.shape)
q_ref, dx_ref, dq_ref, det_ref = rik.rmr(q0, x_des, iters=1)
q=q0.copy(); out=m.ik_rmr(tasks,q,x_des,iters=1,want=("dx","dq","det"))
err=np.abs(out["dq"]-dq_ref).max(axis=1); mag=np.abs(dq_ref).max(axis=1)
order=np.argsort(det_ref)
for i in list(order[:8])+list(order[len(order)//2:len(order)//2+3])+list(order[-3:]):
    print(f"det {det_ref[i]:.3e} ours {out['det'][i]:.3e} |dq| {mag[i]:.3e} err {err[i]:.3e} rel {err[i]/mag[i]:.3e}")
print("quantiles rel err:", np.quantile(err/np.maximum(mag,1e-300),[0.5,0.9,0.99,1.0]))
print("max err for det>1e-4:", err[det_ref>1e-4].max(), " count", (det_ref>1e-4).sum())
print("max err*det:", (err*det_ref).max())

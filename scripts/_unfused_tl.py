import sys,os
sys.path.insert(0,'/root/repo')
import hosnerf_b200.mip360 as M
M.FUSE_IPE = False
sys.argv=['x','2','nerf']
exec(open('/root/repo/scripts/mlp_timeline.py').read())

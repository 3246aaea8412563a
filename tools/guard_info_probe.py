import sys; sys.path.insert(0,'/root/repo')
import hijiki_b200 as hj
ctx=hj.Context(0)
for name,sc in (("cbox",hj.Scene.from_obj('/root/repo/scenes/cbox/cbox.obj')),("cbox_spheres",hj.Scene.from_obj('/root/repo/scenes/cbox/cbox.obj',put_cbox_spheres=True)),("lattice",hj.Scene.spheres(8))):
    ctx.scene_upload(sc.compile()); print(name, ctx.get_info("sphere_guard"))

"""Calibrates the per-element cost of the convert kernel's op classes: homogeneous plans (k identical ops) at
--points, time per launch / k. Used to set the weights of the warp work-item balancer (convert.cu assign_items)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pasture_b200 as pb
from pasture_b200 import PointAttributeDefinition as PAD, PointAttributeDataType as DT, PointLayout, FieldAlignment

ap = argparse.ArgumentParser(); ap.add_argument("--points", type=int, default=50_000_000); args = ap.parse_args()
n = args.points


def timed(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run(name, src_attrs, dst_attrs, packed_src, packed_dst, src_cls, dst_cls, setup=None, fused=False):
    sl = PointLayout(); dl = PointLayout()
    for a in src_attrs: sl.add_attribute(a, FieldAlignment(packed_src))
    for a in dst_attrs: dl.add_attribute(a, FieldAlignment(packed_dst))
    src = src_cls(sl, n, "cuda"); dst = dst_cls(dl, n, "cuda")
    if hasattr(src, "data"): src.data.random_(0, 255)
    else:
        for c in src.columns: c.random_(0, 255)
    cv = pb.BufferLayoutConverter.for_layouts_with_default(sl, dl)
    if setup: setup(cv)
    r = range(0, n)
    mm = torch.zeros(6, dtype=torch.float64, device="cuda")
    fn = (lambda: cv.convert_into_range_with_bounds_device(src, r, dst, r, mm)) if fused else (lambda: cv.convert_into_range(src, r, dst, r))
    ms = timed(fn)
    k = cv.num_mappings()
    bytes_pp = sl.size_of_point_entry() + sum(a.size() for a in dl.attributes())
    print(json.dumps({"case": name, "ms": round(ms, 4), "mappings": k, "ms_per_mapping_per_100M": round(ms / k * 1e8 / n, 4),
                      "GBps": round(bytes_pp * n / ms / 1e6, 1)}), flush=True)


V, H = pb.VectorBuffer, pb.HashMapBuffer
u8 = [PAD(f"a{i}", DT.U8) for i in range(12)]
run("12 x u8 copy, AoS(12 B) -> SoA", u8, u8, 1, 1, V, H)
u16 = [PAD(f"a{i}", DT.U16) for i in range(12)]
run("12 x u16 copy, AoS(24 B) -> SoA", u16, u16, 1, 1, V, H)
def bitfields(cv):
    for i in range(12):
        cv.set_custom_mapping_with_transformation(PAD(f"a{i}", DT.U8), PAD(f"a{i}", DT.U8), pb.ShiftMask(3, 7), True)
run("12 x u8 shift-mask, AoS(12 B) -> SoA", u8, u8, 1, 1, V, H, bitfields)
i32 = [PAD(f"a{i}", DT.I32) for i in range(6)]
f64 = [PAD(f"a{i}", DT.F64) for i in range(6)]
run("6 x i32 -> f64 cast, AoS(24 B) -> SoA", i32, f64, 1, 1, V, H)
def so(cv):
    for i in range(6):
        cv.set_custom_mapping_with_transformation(PAD(f"a{i}", DT.I32), PAD(f"a{i}", DT.F64), pb.ScaleOffset(0.001, 5.0), False)
run("6 x i32 -> f64 scale/offset, AoS(24 B) -> SoA", i32, f64, 1, 1, V, H, so)
pos_i = [PAD("LASLocalPosition", DT.Vec3i32), PAD("pad", DT.U32)]
pos_f = [PAD("Position3D", DT.Vec3f64)]
def pos(cv):
    cv.set_custom_mapping_with_transformation(PAD("LASLocalPosition", DT.Vec3i32), PAD("Position3D", DT.Vec3f64), pb.ScaleOffset(0.001, 5.0), False)
run("Vec3i32 -> Vec3f64 scale/offset (3 ops), AoS(16 B) -> SoA", pos_i, pos_f, 1, 1, V, H, pos)
run("same + fused AABB", pos_i, pos_f, 1, 1, V, H, pos, fused=True)
f64c = [PAD(f"a{i}", DT.F64) for i in range(4)]
run("4 x f64 copy, SoA -> AoS(32 B aligned)", f64c, f64c, 0, 0, H, V)
odd = [PAD("x", DT.U8)] + [PAD(f"a{i}", DT.F64) for i in range(4)]
run("u8 + 4 x f64 copy, SoA -> AoS(33 B packed: byte stores)", odd, odd, 1, 1, H, V)
run("u8 + 4 x f64 copy, AoS(33 B packed: funnel loads) -> SoA", odd, odd, 1, 1, V, H)
def inv(cv):
    for i in range(4):
        cv.set_custom_mapping_with_transformation(PAD(f"a{i}", DT.F64), PAD(f"a{i}", DT.I32), pb.InvScaleOffset(0.001, 5.0), True)
i32d = [PAD(f"a{i}", DT.I32) for i in range(4)]
run("4 x f64 -> i32 inv scale/offset (division), SoA -> SoA", f64c, i32d, 0, 0, H, H, inv)

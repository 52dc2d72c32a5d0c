"""CPU suite: the C-ABI library loads and exports what include/sd_fusion.h declares, struct layouts match the
ctypes mirror, the host-side logic (params, scene, sharding) behaves, and the product never imports the oracle.
No compute call is made here (there is no GPU in this environment)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from semantic_depth_b200 import build, _lib
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from semantic_depth_b200 import _lib
    header = open(os.path.join(ROOT, "include", "sd_fusion.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|void|const char\*)\s+(sd_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sd_fusion.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.sd_abi_version() == 2


def test_struct_layouts_and_defaults(lib):
    from semantic_depth_b200 import _lib
    from semantic_depth_b200.engine import RESULT_DTYPE, params_struct
    from semantic_depth_b200.params import FusionParams
    assert C.sizeof(_lib.SdFrameResult) == RESULT_DTYPE.itemsize == 328
    assert C.sizeof(_lib.SdCamera) == 20 and C.sizeof(_lib.SdPredicate) == 72
    p = _lib.SdParams()
    lib.sd_default_params(C.byref(p), 10.0)
    q = params_struct(FusionParams())
    for name, _ in _lib.SdParams._fields_:
        assert getattr(p, name) == getattr(q, name), name          # the C defaults == the reference literals
    assert p.slab_lo == -((10.0 - 0.02) + 0.05) and p.slab_hi == -((10.0 - 0.02) - 0.05)
    assert lib.sd_fuse_kernel_count(C.byref(p), 0) > 30
    assert lib.sd_ws_bytes(1, 1024, 2048, 0) > 100 << 20 and lib.sd_ws_bytes(1, 8, 6, 0) == 0   # width % 4


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from semantic_depth_b200 import _lib
    from semantic_depth_b200.engine import FusionEngine
    with pytest.raises(_lib.SdError):
        FusionEngine(64, 128)
    import semantic_depth_lib.pcl as pcl
    with pytest.raises((_lib.SdError, RuntimeError, AssertionError)):
        pcl.remove_noise_by_mad(np.zeros((10, 3), np.float32), np.zeros((10, 3), np.uint8), 1, 15.0)


def test_product_never_imports_the_oracle():
    for pkg in ("semantic_depth_b200", "semantic_depth_lib"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{pkg}/{f} imports the oracle"
    code = "import sys; import semantic_depth_b200, semantic_depth_lib.pcl, semantic_depth_b200.stream; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stdout + out.stderr


def test_status_bits_agree_everywhere():
    from oracle import frame_ref
    from semantic_depth_b200 import params
    header = open(os.path.join(ROOT, "include", "sd_fusion.h")).read()
    for name in ("EMPTY_ROAD", "EMPTY_FENCE_LEFT", "EMPTY_FENCE_RIGHT", "MAD_ZERO", "NO_SLAB_POINTS", "SINGULAR_PLANES",
                 "EMPTY_FENCE", "SINGULAR_FIT"):
        shift = int(re.search(rf"SD_ST_{name} = 1 << (\d+)", header).group(1))
        assert getattr(params, f"STATUS_{name}") == 1 << shift == getattr(frame_ref, name)
    assert params.status_to_names(17) == ["EMPTY_ROAD", "NO_SLAB_POINTS"]


def test_scene_is_deterministic_and_shaped():
    from semantic_depth_b200 import scene
    from semantic_depth_b200.params import Intrinsics
    a = scene.make_frame(64, 128, 3)
    b = scene.make_frame(64, 128, 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].shape == (64 * 128, 3) and a[0].dtype == np.float32 and a[1].shape == (2, 64, 128)
    assert not np.array_equal(a[0], scene.make_frame(64, 128, 4)[0])
    label, depth = scene.scene_geometry(256, 512)
    assert set(np.unique(label)) == {0, 1, 2} and depth.max() == 80.0 and depth.min() > 5.0
    q = Intrinsics.cityscapes(512).as_q32()
    assert q.dtype == np.float32 and q[2] == np.float32(-500.0) and q[3] == np.float32(1 / 0.6)
    lg, dp, _ = scene.make_batch(2, 32, 64, first_seed=5)
    assert np.array_equal(lg[1], scene.make_frame(32, 64, 6)[0])
    cloud = scene.make_road_cloud(1000, seed=0)
    assert cloud.shape == (1000, 3) and cloud.dtype == np.float32


def test_shard_frames_partitions_exactly():
    from semantic_depth_b200.stream import shard_frames
    for n in (0, 1, 5, 600, 601, 607):
        for world in (1, 2, 4, 8):
            chunks = [shard_frames(n, r, world) for r in range(world)]
            flat = [i for c in chunks for i in c]
            assert flat == list(range(n))
            assert max(len(c) for c in chunks) - min(len(c) for c in chunks) <= 1
    with pytest.raises(ValueError):
        shard_frames(10, 2, 2)


def _gather_worker(rank, world, n_frames, port, out_dir):
    import torch
    import torch.distributed as dist
    from semantic_depth_b200.stream import gather_results, pack_answers, shard_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_frames(n_frames, rank, world)
    rw = np.array([7.0 + 0.01 * i for i in mine])
    f2f = np.array([7.5 - 0.01 * i for i in mine])
    status = np.array([i % 3 for i in mine])
    allr = gather_results(pack_answers(rw, f2f, status), n_frames)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allr.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 12])
def test_frame_parallel_gather_world2_gloo(tmp_path, n_frames):
    """The N>1 host path: frames sharded over 2 ranks, answers all-gathered into frame order (gloo, CPU)."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n_frames
    mp.spawn(_gather_worker, args=(2, n_frames, port, str(tmp_path)), nprocs=2, join=True)
    expect = np.stack([[7.0 + 0.01 * i for i in range(n_frames)], [7.5 - 0.01 * i for i in range(n_frames)],
                       [float(i % 3) for i in range(n_frames)]], axis=1)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npy"))
        np.testing.assert_array_equal(got, expect)


def test_bench_alg_bytes_matches_survey_formula():
    sys.path.insert(0, ROOT)
    import bench
    counts = {"road_gather": 606837, "fence_gather": 691785, "road_z": 452841, "road_mad_y": 450831, "road_mad_x": 450831,
              "road_plane": 450831, "road_sor": 391988, "road_ror": 390743, "fence_mad_y": 691424, "fence_abs_z": 666758,
              "left_split": 326266, "right_split": 340492, "left_mad_x": 325826, "left_plane": 325826,
              "right_mad_x": 232772, "right_plane": 232772}
    b = bench.b_alg_bytes(counts, 1024 * 2048)
    assert abs(b / 1e6 - 203.9) < 0.5      # 203.9 MB for the seed-0 frame


def test_header_is_plain_c_and_layouts_match_field_by_field(tmp_path):
    """include/sd_fusion.h compiles as C99 (no C++ / CUDA / torch types at the boundary) and every field of every struct
    sits where the ctypes mirror puts it."""
    import shutil
    from semantic_depth_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"SdCamera": _lib.SdCamera, "SdParams": _lib.SdParams, "SdFrameResult": _lib.SdFrameResult,
               "SdPredicate": _lib.SdPredicate, "SdFcnHeadWeights": _lib.SdFcnHeadWeights}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sd_fusion.h"', "int main(void) {"]
    for cname, ctype in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ctype._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("SD_NUM_COUNTS %d\\n", (int)SD_NUM_COUNTS);', "  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, ctype in structs.items():
        assert int(got[cname]) == C.sizeof(ctype), cname
        for fname, _ in ctype._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ctype, fname).offset, f"{cname}.{fname}"
    assert int(got["SD_NUM_COUNTS"]) == _lib.SD_NUM_COUNTS


def test_frame_processor_constructor_contract():
    from semantic_depth_b200.frame_processor import FrameProcessor

    class Seg:
        def logits(self, frame):
            return None

    class Dep:
        def disparities(self, frame):
            return None

    p = FrameProcessor(Seg(), Dep(), (256, 512), approach="rw", depth=12.5)
    assert p.params.approach == "rw" and p.params.depth == 12.5 and p.input_shape == (256, 512)
    assert p._intrinsics(2048).disparity_mult == 2048.0                  # semantic_depth.py:109: the original width
    assert FrameProcessor(Seg(), Dep(), (256, 512), disp_multiplier=3800)._intrinsics(2048).disparity_mult == 3800.0   # sequence:105
    with pytest.raises(TypeError):
        FrameProcessor(object(), Dep(), (256, 512))
    with pytest.raises(ValueError):
        FrameProcessor(Seg(), Dep(), (256, 512), approach="f2f")


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line with the contract keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--height", "64", "--width", "128", "--cpu-procs", "2"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fusion_frames_per_sec_1024x2048" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and 2 <= d["cpu_baseline"]["cores"] <= os.cpu_count()      # 2 processes x their k-d tree threads
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_host_placement_logic_on_cpu():
    """semantic_depth_b200.hostmem: the pieces that do not need a GPU -- CPU-list parsing, the candidate rank -> GPU maps of
    a job that is smaller than the box, placement calls degrading to no-ops where sysfs / the device is missing."""
    import numpy as np
    import torch
    from semantic_depth_b200 import hostmem as h
    assert h._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and h._parse_cpulist("") == []
    assert h.candidate_device_maps(8, 8) == {"first": list(range(8))}
    assert h.candidate_device_maps(4, 8) == {"first": [0, 1, 2, 3], "last": [4, 5, 6, 7], "spread": [0, 2, 4, 6]}
    assert h.candidate_device_maps(2, 8)["spread"] == [0, 4] and h.candidate_device_maps(3, 4)["last"] == [1, 2, 3]
    info = h.gpu_locality(0)                                  # no driver here: reported, not raised
    assert info["numa_node"] == -1 and info["local_cpus"] == []
    done = h.bind_to_gpu(0)
    assert done["cpus"] is None and done["mempolicy"] is False
    t = torch.zeros(1 << 20, dtype=torch.uint8)
    hist = h.node_histogram(t.data_ptr(), t.numel())
    assert sum(hist.values()) > 0 and all(isinstance(k, int) for k in hist)
    # one rank, or as many ranks as GPUs: the plain map, nothing is measured
    dev, rep = h.choose_device(0, 1)
    assert dev == 0 and rep["map"] == "first"


def test_torch_op_layer_loads_and_rejects_host_tensors():
    """the TORCH_LIBRARY layer over the C ABI checks device / dtype / contiguity in C++ and raises RuntimeError."""
    import ctypes as C
    import torch
    from semantic_depth_b200 import _lib
    ops = _lib.load_ops()
    assert int(ops.abi_version()) == 2 and int(ops.result_bytes()) == C.sizeof(_lib.SdFrameResult)
    cam = _lib.struct_tensor(_lib.SdCamera())
    ps = _lib.struct_tensor(_lib.SdParams())
    assert cam.numel() == C.sizeof(_lib.SdCamera) and ps.numel() == C.sizeof(_lib.SdParams)
    res = torch.zeros(C.sizeof(_lib.SdFrameResult), dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ops.fuse_frames(torch.zeros(1, 8, 3), torch.zeros(1, 2, 2, 4), cam, ps, None, None, None, 1, res)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ops.fuse_frames_scores(torch.zeros(1, 1, 1, 3), torch.zeros(16, 16, 3, 3), torch.zeros(3), torch.zeros(1, 2, 8, 8), cam, ps, 1, res)

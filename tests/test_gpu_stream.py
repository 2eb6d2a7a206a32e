"""K2 on a real B200 against frames the UNMODIFIED reference returned while its GL stream was recorded and rasterised
(tests/golden/stream_*.npz; oracle/gl_record.py + oracle/gen_stream_golden.py): every configured level and every other
reference level, frames along rollouts (incl. Maze with domain randomisation, the PickupObjects / CollectHealth frames
that still show the just-picked-up object, 160 x 120), depth maps, map views and occlusion-query sets."""
import pytest

from helpers import stream_cases, stream_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", stream_cases())
def test_k2_matches_reference_stream_frames(libmwb_path, name):
    st = stream_parity(name, libmwb_path)
    print("%s: %d frames (%d with a just-removed object), worst %d LSB, %.4f%% of channel values identical, %d map views, "
          "%d visibility sets" % (name, st["frames"], st["events"], st["worst"], 100.0 * st["same"] / st["total"],
                                  st["tops"], st["vis"]))
    assert st["frames"] >= 4 and st["worst"] <= 1 and st["same"] / st["total"] > 0.995
    assert st["cams"] > 0 and st["cam_exact"] / st["cams"] > 0.98          # camera vs Agent.cam_pos / cam_dir / cam_fov_y

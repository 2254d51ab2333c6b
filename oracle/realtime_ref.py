"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the per-frame arithmetic of the reference's realtime script
(src/openpose_3dpose_sandbox_realtime.py:69-171,178-195).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; the product path (p3d.realtime -> libp3d.so) never does.

Parity pinned: oracle/make_golden_realtime.py EXECUTES the reference's own source lines (:69-163 and :178-195, cut
out of the file by line number - the script as a whole needs TensorFlow/OpenCV and cannot be imported) on seeded
synthetic keypoints and writes tests/golden/realtime.npz; tests/test_oracle_realtime.py holds this restatement to it
bit for bit.  The un-normalisation step is geometry_ref.unnormalize (pinned on data_utils.unNormalizeData)."""
import numpy as np

ORDER = [15, 12, 25, 26, 27, 17, 18, 19, 1, 2, 3, 6, 7, 8]   # :20


def keypoints_to_xy(pose_keypoints_2d):
    """:69-135 - confidence stripping (:71-79) and the BODY_25 -> COCO list surgery (:86-135), written as the
    reference writes it (del + overwrite), not as its net effect."""
    _data = list(pose_keypoints_2d)
    xy = []
    if len(_data) >= 53:
        for o in range(0, len(_data), 3):
            xy.append(_data[o])
            xy.append(_data[o + 1])
    else:
        xy = _data
    if len(xy) > 54:
        _xy = xy[0:19 * 2]
        for x in range(len(xy)):
            if x == 8 * 2:
                del _xy[x]
            if x == 8 * 2 + 1:
                del _xy[x]
            for src in range(9, 19):                      # "map jnt src to src-1" (:95-133)
                if x == src * 2:
                    _xy[(src - 1) * 2] = xy[x]
                    _xy[(src - 1) * 2 + 1] = xy[x + 1]
        xy = _xy
    return xy


def frontend(xy, data_mean_2d, data_std_2d, dim_to_use_2d, enc_prev=None):
    """:137-163.  enc_prev: the [1,64] array the reference carries from the previous frame (zeros on the first one, :28-29).
    Returns (enc_in [1,32] float64, spine_x, spine_y)."""
    enc_in = np.zeros((1, 64)) if enc_prev is None else np.array(enc_prev, dtype=np.float64, copy=True)
    joints_array = np.zeros((1, 36))
    for o in range(36):
        joints_array[0][o] = xy[o]
    _data = joints_array[0]
    for i in range(len(ORDER)):
        for j in range(2):
            enc_in[0][ORDER[i] * 2 + j] = _data[i * 2 + j]
    for j in range(2):
        enc_in[0][0 * 2 + j] = (enc_in[0][1 * 2 + j] + enc_in[0][6 * 2 + j]) / 2        # Hip
        enc_in[0][14 * 2 + j] = (enc_in[0][15 * 2 + j] + enc_in[0][12 * 2 + j]) / 2     # Neck/Nose
        enc_in[0][13 * 2 + j] = 2 * enc_in[0][12 * 2 + j] - enc_in[0][14 * 2 + j]       # Thorax
    spine_x, spine_y = enc_in[0][24], enc_in[0][25]
    enc_in = enc_in[:, dim_to_use_2d]
    enc_in = np.divide((enc_in - data_mean_2d[dim_to_use_2d]), data_std_2d[dim_to_use_2d])
    return enc_in, spine_x, spine_y


def display_transform(poses3d, spine_x, spine_y):
    """:178-195 (the loops, as written)."""
    poses3d = np.array(poses3d, dtype=np.float64, copy=True)
    _max, _min = 0, 10000
    for i in range(poses3d.shape[0]):
        for j in range(32):
            tmp = poses3d[i][j * 3 + 2]
            poses3d[i][j * 3 + 2] = poses3d[i][j * 3 + 1]
            poses3d[i][j * 3 + 1] = tmp
            if poses3d[i][j * 3 + 2] > _max:
                _max = poses3d[i][j * 3 + 2]
            if poses3d[i][j * 3 + 2] < _min:
                _min = poses3d[i][j * 3 + 2]
    for i in range(poses3d.shape[0]):
        for j in range(32):
            poses3d[i][j * 3 + 2] = _max - poses3d[i][j * 3 + 2] + _min
            poses3d[i][j * 3] += (spine_x - 630)
            poses3d[i][j * 3 + 2] += (500 - spine_y)
    return poses3d

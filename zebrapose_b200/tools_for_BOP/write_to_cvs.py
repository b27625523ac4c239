"""Drop-in for zebrapose/tools_for_BOP/write_to_cvs.py:6-62 (called at test.py:506, test_vivo.py:190): the BOP result
file `scene_id,im_id,obj_id,score,R,t,time`.  Same signature and byte-identical output: one row per estimate whose score
is not -1, R as 9 and t as 3 space-separated str() values (row-major; t in mm), time fixed to -1."""
import os

import numpy as np

HEADER = "scene_id,im_id,obj_id,score,R,t,time\n"


def format_row(scene_id, img_id, obj_id, score, r, t):
    rr = [str(r[i][j]) for i in range(3) for j in range(3)]
    tt = [str(t[i][0]) for i in range(3)]
    return "%s,%s,%s,%s,%s,%s,-1\n" % (str(scene_id), str(img_id), str(obj_id), str(score), " ".join(rr), " ".join(tt))


def write_cvs(evaluation_result_path, filename, obj_id, scene_id_, img_id_, r_, t_, scores):
    path = os.path.join(evaluation_result_path, filename + ".csv")
    rows = [HEADER]
    for scene_id, img_id, r, t, score in zip(scene_id_, img_id_, r_, t_, scores):
        if score == -1:
            continue
        rows.append(format_row(scene_id, img_id, obj_id, score, r, t))
    with open(path, "w") as f:
        f.write("".join(rows))


def write_batch(evaluation_result_path, filename, obj_id, scene_ids, img_ids, poses12, scores):
    """Same file from the batched engine output: poses12 float64 [B,12] (R row-major | t), e.g. the gathered
    records of sharding.gather_poses()."""
    p = np.asarray(poses12, np.float64).reshape(-1, 12)
    write_cvs(evaluation_result_path, filename, obj_id, scene_ids, img_ids,
              [q[:9].reshape(3, 3) for q in p], [q[9:].reshape(3, 1) for q in p], scores)

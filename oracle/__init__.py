"""CPU oracle for the lidar voxelization / BEV-rasterization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg (``cpu_baseline`` / ``--impl reference``) may import it.  The product
package (``lyft-3d-object-detection_b200/``) never imports this package and
fails loudly when its CUDA library is missing.

Contents (each function cites the reference file:line it restates):

* ``bev_oracle``    - numpy restatement of the BEV rasteriser closures
                      (generating-dataset/generating_train_bev.py:47-104) and of
                      the sensor->car transform (lyft_dataset_sdk/utils/data_classes.py:188-195).
                      PINNED: restated line for line from code that is in the
                      reference tree; checked against the survey's known answers
                      on the bundled sweep.
* ``voxel_oracle``  - hard voxelizer (spconv ``VoxelGeneratorV2`` contract).
                      PARITY UNPINNED at the spconv boundary: spconv is an
                      un-vendored, un-pinned third-party dependency whose source
                      is not under /root/reference.  The
                      restatement follows the in-tree sibling
                      second/second/utils/simplevis.py:9-61 and is pinned against
                      THAT function (executed from /root/reference when the
                      golden vectors are generated - oracle/gen_golden.py).
* ``pillar_oracle`` - PillarFeatureNet decoration + PointPillarsScatter
                      (second/second/pytorch/models/pointpillars.py).  PINNED:
                      golden vectors are produced by executing the reference's own
                      file (oracle/gen_golden.py + oracle/ref_loader.py).
"""

"""CPU oracle for the loop-closure hot path: a float64 NumPy restatement of the reference's algorithms.

TEST INFRASTRUCTURE ONLY. Nothing under ``deeploopcloser_b200/`` (the product) may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and
there only as the checker / the timed CPU baseline.

Pinning status (see DESIGN.md "Oracle"):
  * patches, similarity, hamming, compressed_size: PINNED against the reference's own code, imported from
    /root/reference in the authoring container by ``tests/golden/make_golden.py`` (fixtures committed under
    ``tests/golden/``), and against the reference's only known-answer test (test/TensorflowWrapperTest.py:11-21).
  * sda (encoder forward) and cnnvtl (conv head): PARITY UNPINNED - the reference classes need TensorFlow 1.x, which
    is not installable here, and the reference ships neither trained SDA weights nor the AlexNet blob (git-LFS
    pointer). The restatement follows the graph definitions line by line (citations in each function).
  * matcher (cosine / L2 / dot top-k, threshold): PARITY UNPINNED by construction - the reference has no such step;
    the oracle is the definition.
"""

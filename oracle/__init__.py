"""CPU oracle (test infrastructure only; never imported by faster_rcnn_b200)."""

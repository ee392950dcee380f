"""`turtleFSI` package name for problem files that import `turtleFSI.problems` / `turtleFSI.modules`
(turtleFSI/problems/TF_fsi.py:19-20): everything resolves to turtlefsi_b200."""

"""Copies the UNMODIFIED reference checkout into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The reference is pure Python with no setup.py / pyproject, so "installing" it is a copy:
  baseline/_ref/tamago     /root/reference as it is                       (9x9)
  baseline/_ref/tamago19   the same tree with board/constant.py:4 set to BOARD_SIZE = 19 -- the reference's own way of
                           selecting 19x19 ("Board size is a source-level constant"); nothing else differs.
Called by __graft_entry__.build() when /root/reference exists (the build container); the GPU box only uses the copy.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def install(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(REF, "tamago"))
    for name, size in (("tamago", 9), ("tamago19", 19)):
        dst = os.path.join(REF, name)
        if os.path.isdir(dst) and not force:
            continue
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(SRC, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "img", "doc", "*.bin"))
        os.system(f"chmod -R u+w {dst}")
        if size != 9:
            p = os.path.join(dst, "board", "constant.py")
            s = open(p).read()
            assert "BOARD_SIZE = 9" in s
            open(p, "w").write(s.replace("BOARD_SIZE = 9", f"BOARD_SIZE = {size}"))
    return True


def ref_dir(size):
    d = os.path.join(REF, "tamago" if size <= 9 else "tamago19")
    return d if os.path.isdir(d) else None


if __name__ == "__main__":
    print(install(force=True))

"""TEST INFRASTRUCTURE.  Writes a patched copy of the reference's db_construction.cpp into the (git-ignored) build
directory: the two functions INTEGRATION.md replaces are swapped for oracle/integration/*.inc, nothing else changes.

  python apply_patch.py /root/reference/src/db_construction.cpp oracle/_ref/patched/db_construction.cpp
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def replace_function(src: str, head: str, body: str) -> str:
    a = src.index(head)
    i = src.index("{", a)
    depth = 0
    while True:
        if src[i] == "{":
            depth += 1
        elif src[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return src[:a] + body.rstrip() + "\n" + src[i + 1:]


def main():
    src = open(sys.argv[1]).read()
    src = replace_function(src, "void DbConstruction::CalculateAccessibility(",
                           open(os.path.join(HERE, "CalculateAccessibility.inc")).read())
    src = replace_function(src, "void DbConstruction::ConstructSuffixArray(",
                           open(os.path.join(HERE, "ConstructSuffixArray.inc")).read())
    src = '#include "priblast_acc.h"\n#include <cstdlib>\n#include <fstream>\n' + src
    os.makedirs(os.path.dirname(os.path.abspath(sys.argv[2])), exist_ok=True)
    open(sys.argv[2], "w").write(src)


if __name__ == "__main__":
    main()

"""oracle/ref_extract.py — TEST INFRASTRUCTURE ONLY (used by `make -C oracle ref`).
Copies whole function definitions VERBATIM out of a reference source file into a build intermediate under oracle/_ref/
(git-ignored, never committed), so that functions of a translation unit that cannot be compiled as a whole here
(FullSystem/CoarseTracker.cpp needs OpenCV, PCL, Sophus and most of DSO) can still be compiled and run as the reference
wrote them: ref_tracker.cpp #includes the intermediate inside `namespace dso`.
A definition starts at the line that begins with the given signature prefix (leading whitespace included) and ends at the
first following line that is exactly `}` at the same indentation (the reference closes every function that way).
A signature written as `<prefix>@-N` also takes the N lines before it (Sophus puts `inline static` on its own line).
usage: ref_extract.py <source> <out.inc> <signature prefix> [<signature prefix> ...]
       ref_extract.py --defines <header> <out.inc> <macro prefix>      (copies `#define <prefix>...` lines)"""
import sys


def main():
    if sys.argv[1] == "--defines":
        src, out, prefix = sys.argv[2:5]
        lines = [l for l in open(src, encoding="utf-8", errors="replace") if l.lstrip().startswith("#define " + prefix)]
        assert lines, f"no #define {prefix}* in {src}"
        open(out, "w").write("".join(lines))
        return
    src, out, sigs = sys.argv[1], sys.argv[2], sys.argv[3:]
    text = open(src, encoding="utf-8", errors="replace").read().split("\n")
    chunks = []
    for sig in sigs:
        before = 0
        if "@-" in sig and sig.rsplit("@-", 1)[1].isdigit():
            sig, nb = sig.rsplit("@-", 1)
            before = int(nb)
        starts = [i for i, l in enumerate(text) if l.startswith(sig)]
        assert len(starts) == 1, f"{sig!r}: {len(starts)} definitions in {src}"
        i = starts[0]
        indent = text[i][: len(text[i]) - len(text[i].lstrip())]  # a member function defined inside a class ends at `<indent>}`
        j = i
        while text[j].rstrip("\r\n ") != indent + "}":
            j += 1
        i -= before
        chunks.append(f"// ---- {src}:{i + 1}-{j + 1} (verbatim)\n" + "\n".join(text[i : j + 1]) + "\n")
    open(out, "w").write("\n".join(chunks))


if __name__ == "__main__":
    main()

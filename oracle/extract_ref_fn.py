"""TEST INFRASTRUCTURE (build step of oracle/_ref): print the text of one member-function DEFINITION of a reference source file,
found by its signature prefix, from the line of the signature to the closing brace that ends the function body.  The text goes
straight into a file under oracle/_ref/ (git-ignored) that oracle/ref_match.cpp includes between its stand-in declarations, so the
reference's own statements are what gets compiled; nothing of it is stored in the repository.
usage: python extract_ref_fn.py <reference file> <signature prefix, e.g. 'void Frame::ComputeStereoMatches()'> [more prefixes ...]"""
import sys


def extract(text: str, prefix: str) -> str:
    at = text.find("\n" + prefix)
    if at < 0:
        raise SystemExit(f"signature not found: {prefix}")
    at += 1
    i = text.index("{", at)
    depth, j = 0, i
    in_line_comment = in_block_comment = in_str = in_chr = False
    while j < len(text):
        c, n = text[j], text[j + 1] if j + 1 < len(text) else ""
        if in_line_comment:
            in_line_comment = c != "\n"
        elif in_block_comment:
            if c == "*" and n == "/":
                in_block_comment = False; j += 1
        elif in_str:
            if c == "\\": j += 1
            elif c == '"': in_str = False
        elif in_chr:
            if c == "\\": j += 1
            elif c == "'": in_chr = False
        elif c == "/" and n == "/": in_line_comment = True
        elif c == "/" and n == "*": in_block_comment = True
        elif c == '"': in_str = True
        elif c == "'": in_chr = True
        elif c == "{": depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[at:j + 1] + "\n"
        j += 1
    raise SystemExit(f"unbalanced braces after: {prefix}")


if __name__ == "__main__":
    src = open(sys.argv[1], encoding="utf-8", errors="replace").read()
    for p in sys.argv[2:]:
        sys.stdout.write(f"// ---- {sys.argv[1]}: {p}\n")
        sys.stdout.write(extract(src, p))

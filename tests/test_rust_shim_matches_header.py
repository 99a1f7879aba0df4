"""The Rust host side (rust/) cannot be compiled in this image; keep it honest against the C header mechanically:
every PB200_API function is declared in rust/src/sys.rs's extern block, every header struct has a #[repr(C)] mirror
with the same field names in the same order, and the constants agree."""
import os
import re

from conftest import ROOT

HEADER = open(os.path.join(ROOT, "include", "phonic_b200.h")).read()
SYS = open(os.path.join(ROOT, "rust", "src", "sys.rs")).read()
LIB = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()


def test_every_entry_point_is_bound():
    declared = sorted(set(re.findall(r"PB200_API[^;(]*?\b(pb200_\w+)\s*\(", HEADER)))
    bound = sorted(set(re.findall(r"pub fn (pb200_\w+)\s*\(", SYS)))
    assert declared == bound


def c_struct_fields(name):
    m = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (name, name), HEADER, re.S)
    assert m, name
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith(("pb200_", "struct")) else decl.split(None, 1)[1]
        for n in names.split(","):
            fields.append(re.sub(r"\[.*\]", "", n).strip().lstrip("*"))
    return fields


def rust_struct_fields(name):
    m = re.search(r"pub struct %s\s*\{(.*?)\n\}" % name, SYS, re.S)
    assert m, name
    return re.findall(r"pub (\w+):", m.group(1))


def test_struct_mirrors_have_the_same_fields():
    for name in re.findall(r"typedef struct (pb200_\w+)\s*\{", HEADER):
        assert c_struct_fields(name) == rust_struct_fields(name), name


def test_constants_agree():
    for name, value in re.findall(r"\b(PB200_(?:EV|FX|ERR|MSG|EVF|MOVE)_\w+)\s*=\s*(\d+)", HEADER) + re.findall(r"#define (PB200_(?:EVF|MSG)_\w+) (\d+)u", HEADER):
        m = re.search(r"pub const %s: \w+ = (\d+);" % name, SYS)
        assert m and int(m.group(1)) == int(value), name


def test_player_mirror_covers_the_player_methods():
    for method in ["play_file", "play_file_buffer", "add_sampler", "remove_generator", "add_mixer", "remove_mixer", "add_effect", "move_effect",
                   "remove_effect", "stop_all_sources", "render_to_wav", "output_sample_frame_position"]:
        assert re.search(r"pub fn %s\b" % method, LIB), method
    assert "impl OutputDevice for B200Output" in LIB

"""Stand-in for gridtools_cpp (headers absent; only non-numpy backends need them)."""


def get_include_dir():
    return "/nonexistent"


def get_cmake_dir():
    return "/nonexistent"

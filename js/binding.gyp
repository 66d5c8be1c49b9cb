{
  "targets": [{
    "target_name": "meyda_b200",
    "sources": ["addon.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../meyda_b200/_lib", "-lmeyda_b200", "-Wl,-rpath,<(module_root_dir)/../meyda_b200/_lib"],
    "cflags_cc": ["-std=c++17"]
  }]
}

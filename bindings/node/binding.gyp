{
  "targets": [{
    "target_name": "tendrils_b200",
    "sources": ["tendrils_b200_napi.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../tendrils_b200/lib", "-ltendrils_b200",
                  "-Wl,-rpath,<(module_root_dir)/../../tendrils_b200/lib"]
  }]
}

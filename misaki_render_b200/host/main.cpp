// misaki_b200 -- command-line renderer, the counterpart of reference src/apps/main.cpp (which hard-codes its
// scene path, main.cpp:66): misaki_b200 <scene.xml> [-o output] [-D name=value ...] [-q]
#include "render.h"

#include <cstdio>
#include <cstring>

using namespace misaki;

int main(int argc, char **argv) {
    std::string scene_path, output;
    xml::ParameterList params;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-o") && i + 1 < argc) output = argv[++i];
        else if (!strcmp(argv[i], "-q")) set_log_level(Warn);
        else if (!strncmp(argv[i], "-D", 2)) {
            std::string kv = argv[i][2] ? argv[i] + 2 : (i + 1 < argc ? argv[++i] : "");
            size_t eq = kv.find('=');
            if (eq == std::string::npos) { fprintf(stderr, "-D expects name=value\n"); return 2; }
            params.emplace_back(kv.substr(0, eq), kv.substr(eq + 1));
        } else scene_path = argv[i];
    }
    if (scene_path.empty()) {
        fprintf(stderr, "usage: %s <scene.xml> [-o output.exr] [-D name=value] [-q]\n", argv[0]);
        return 2;
    }
    try {
        std::string exe = argv[0];
        size_t slash = exe.find_last_of('/');
        std::string exe_dir = slash == std::string::npos ? "." : exe.substr(0, slash);
        get_file_resolver()->append(exe_dir); // main.cpp:67
        get_file_resolver()->append(exe_dir + "/.."); // lib/ sits next to data/ in the package
        slash = scene_path.find_last_of('/');
        get_file_resolver()->prepend(slash == std::string::npos ? "." : scene_path.substr(0, slash)); // main.cpp:68
        ref<Object> root = xml::load_file(scene_path, params);
        Scene *scene = dynamic_cast<Scene *>(root.get());
        if (!scene) Throw("Root element of the input file must be a <scene> tag!");
        Sensor *sensor = scene->sensor();
        if (!sensor) Throw("The scene has no sensor");
        Film *film = sensor->film();
        film->set_destination_file(output.empty() ? scene_path : output);
        if (!scene->integrator()) Throw("No integrator specified for scene");
        if (scene->integrator()->render(scene, sensor)) film->develop();
        else { Log(Warn, "Rendering failed, result not saved."); return 3; } // (main.cpp:38-43 only warns; scripts need the status)
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}

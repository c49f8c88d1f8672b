// A Stage 6 style application compiled against the drop-in headers with
// -DRAYITO_B200_STAGE=6 ("Stage 6 applications").  It only builds and
// prepares the scene (host work); rendering needs a GPU and is covered by the -m gpu tests.
#include <cstdio>
#include <list>

#include "rayito.h"
#include "RMesh.h"
#include "scene_recipes_s6.h"

int main(int argc, char** argv)
{
    if (argc < 2)
        return 2;
    Rayito::ShapeSet masterSet;
    rayito_recipes::SceneStore store;
    if (!rayito_recipes::buildStage6Scene(masterSet, store, argv[1]))
        return 3;
    // Stage 6 constructor signature (no shutter)
    Rayito::PerspectiveCamera cam(30.0f, Rayito::Point(-2.0f, 5.0f, 15.0f), Rayito::Point(0.0f, 0.0f, 0.0f),
                                  Rayito::Point(0.0f, 1.0f, 0.0f), 16.0f, 0.0f);
    std::vector<Rayito::Shape*> lights;
    masterSet.findLights(lights);
    masterSet.prepare();
    rayito_b200::FlatScene flat;
    flat.semantics = rayito_b200::stageSemantics();
    if (!masterSet.flattenScene(flat, lights))
        return 4;
    RtCamera rc;
    cam.describe(rc);
    std::printf("semantics=%u shapes=%zu lights=%zu top_nodes=%zu shutter=%g,%g\n", flat.semantics,
                flat.shapes.size(), flat.lights.size(), flat.topNodes.size(), rc.shutter_open, rc.shutter_close);
    return 0;
}

// Ownership and camera description shared by the scene recipes
// (scene_recipes.h: Stage 7 scenes, scene_recipes_s6.h: the Stage 6 scene).
// API-neutral like the recipes themselves: include AFTER the "rayito.h" and
// "RMesh.h" of whichever implementation is in use.
#ifndef RAYITO_B200_SCENE_STORE_H
#define RAYITO_B200_SCENE_STORE_H

#include <cstddef>
#include <vector>

// A build may define RAYITO_RECIPE_WRAP_MESH(ptr) to substitute a Mesh subclass
// (the oracle uses this to record which face/triangle won a hit).
#ifndef RAYITO_RECIPE_WRAP_MESH
#define RAYITO_RECIPE_WRAP_MESH(meshPtr) (meshPtr)
#endif

namespace rayito_recipes
{

// Owns every heap object of a recipe scene (the reference keeps them on the GUI
// thread's stack; a library needs them to outlive the call that built them).
struct SceneStore
{
    std::vector<Rayito::Material*> materials;
    std::vector<Rayito::Shape*> shapes;
    // Shapes wrapped by a ShapeLight; not members of the set themselves
    std::vector<Rayito::Shape*> wrapped;

    SceneStore() { }
    ~SceneStore()
    {
        for (size_t i = 0; i < shapes.size(); ++i) delete shapes[i];
        for (size_t i = 0; i < wrapped.size(); ++i) delete wrapped[i];
        for (size_t i = 0; i < materials.size(); ++i) delete materials[i];
    }

    template <typename M> M* keep(M* m) { materials.push_back(m); return m; }
    template <typename S> S* add(S* s) { shapes.push_back(s); return s; }
    template <typename S> S* hold(S* s) { wrapped.push_back(s); return s; }

private:
    SceneStore(const SceneStore&);
    SceneStore& operator=(const SceneStore&);
};

struct CameraSpec
{
    float fov;
    float origin[3];
    float target[3];
    float up[3];
    float focalDistance, lensRadius, shutterOpen, shutterClose;
};

} // namespace rayito_recipes

#endif // RAYITO_B200_SCENE_STORE_H

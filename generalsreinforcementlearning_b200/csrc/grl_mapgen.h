// grl_mapgen.h — host map generation (mapgen/generator.go restated; see grl_mapgen.cpp)
#pragma once
#include <cstdint>
#include <utility>

namespace grl {

struct MapParams {  // mapgen.MapConfig, generator.go:12-22
  int players;
  int city_ratio;
  int city_start_army;
  int spacing;
  int veins;
  int min_vein;
  int max_vein;
};

MapParams DefaultMapParams(int w, int h, int players, int city_ratio, int city_start_army, int min_general_spacing);

// Generator.GenerateMap with rand.New(rand.NewSource(seed)); planes are int32[w*h].
// Returns false when a general cannot be placed (generator.go:252).
bool GenerateMap(int w, int h, const MapParams &mp, int64_t seed, int32_t *owner, int32_t *army, int32_t *type);

}  // namespace grl

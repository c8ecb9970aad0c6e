// The layer table of the KFNet hot path, shared by the inference schedules (kfn_api.cu) and the training path (train.cu).
// cnn_wrapper/SCoordNet.py:21-32, KFNet/KFNet.py:318-338, cnn_wrapper/OFlowNet.py:19-41,50-55
#pragma once
#include <string>

namespace kfn {

enum Kind { CONV = 0, DECONV = 1, DENSE = 2 };
struct LayerSpec { const char* scope; const char* name; Kind kind; int k, cin, cout, stride; bool relu; };

static const LayerSpec kLayers[] = {
    {"ScoreNet", "conv1a", CONV, 3, 3, 64, 1, true},      {"ScoreNet", "conv1b", CONV, 3, 64, 64, 1, true},
    {"ScoreNet", "conv2a", CONV, 3, 64, 256, 2, true},    {"ScoreNet", "conv2b", CONV, 3, 256, 256, 1, true},
    {"ScoreNet", "conv3a", CONV, 3, 256, 512, 2, true},   {"ScoreNet", "conv3b", CONV, 3, 512, 512, 1, true},
    {"ScoreNet", "conv4a", CONV, 3, 512, 1024, 2, true},  {"ScoreNet", "conv4b", CONV, 3, 1024, 1024, 1, true},
    {"ScoreNet", "conv5", CONV, 3, 1024, 512, 1, true},   {"ScoreNet", "conv6", CONV, 3, 512, 256, 1, true},
    {"ScoreNet", "conv7", CONV, 1, 256, 128, 1, true},    {"ScoreNet", "prediction", CONV, 1, 128, 4, 1, false},
    {"Temporal", "feat1", CONV, 3, 3, 16, 1, true},       {"Temporal", "feat2", CONV, 3, 16, 32, 2, true},
    {"Temporal", "feat3", CONV, 3, 32, 32, 1, true},      {"Temporal", "feat4", CONV, 3, 32, 64, 2, true},
    {"Temporal", "feat5", CONV, 3, 64, 64, 1, true},      {"Temporal", "feat6", CONV, 3, 64, 128, 2, true},
    {"Temporal", "feat7", CONV, 3, 128, 32, 1, false},
    {"Temporal", "conv0", CONV, 3, 32, 32, 1, true},      {"Temporal", "conv1a", CONV, 3, 32, 32, 2, true},
    {"Temporal", "conv1b", CONV, 3, 32, 32, 1, true},     {"Temporal", "conv2a", CONV, 3, 32, 64, 2, true},
    {"Temporal", "conv2b", CONV, 3, 64, 64, 1, true},     {"Temporal", "conv3a", CONV, 3, 64, 128, 2, true},
    {"Temporal", "conv3b", CONV, 3, 128, 128, 1, true},   {"Temporal", "upconv2", DECONV, 3, 128, 64, 2, true},
    {"Temporal", "conv4", CONV, 3, 128, 64, 1, true},     {"Temporal", "upconv1", DECONV, 3, 64, 32, 2, true},
    {"Temporal", "conv5", CONV, 3, 64, 32, 1, true},      {"Temporal", "upconv0", DECONV, 3, 32, 16, 2, true},
    {"Temporal", "conv6", CONV, 3, 48, 16, 1, true},      {"Temporal", "prediction", CONV, 3, 16, 1, 1, false},
    {"Temporal", "fc1", DENSE, 1, 128, 64, 1, true},      {"Temporal", "fc2", DENSE, 1, 64, 32, 1, true},
    {"Temporal", "uncertainty", DENSE, 1, 32, 1, 1, false},
};
constexpr int kNumLayers = sizeof(kLayers) / sizeof(kLayers[0]);

inline int layer_index(const std::string& scope_name) {
  for (int i = 0; i < kNumLayers; ++i)
    if (scope_name == std::string(kLayers[i].scope) + "/" + kLayers[i].name) return i;
  return -1;
}
inline const LayerSpec* find_layer(const std::string& scope_name) {
  const int i = layer_index(scope_name);
  return i >= 0 ? &kLayers[i] : nullptr;
}
inline int64_t layer_kernel_elems(const LayerSpec& L) { return (int64_t)(L.kind == DENSE ? 1 : L.k * L.k) * L.cin * L.cout; }

}  // namespace kfn

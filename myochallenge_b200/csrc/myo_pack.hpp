// Host-side packer: turns the parsed MJB model (+ task configuration) into the flat device tables
// the step kernel reads (DevModel) and derives the tree / sparsity metadata the kernel phases use.
#pragma once
#include <string>
#include <vector>

#include "myo_dev.hpp"
#include "myo_model.hpp"

namespace myo {

struct PackedModel {
  DevModel dm{};                 // tables are word offsets into the table block; scratch layout with the FAST capacities
  DevModel dm_full{};            // same tables, scratch layout with MuJoCo's own capacities (nconmax contacts, njmax rows)
  std::vector<int> ibuf;         // all int tables, concatenated
  std::vector<float> fbuf;       // all float tables, concatenated
  std::vector<TabF*> ffix;       // float tables: offsets get shifted by ibuf.size() once packing is done
  std::vector<float> tables;     // ibuf (bit-cast) followed by fbuf: the block staged into shared memory
  float* d_tables = nullptr;
  int lanes = 32;                // tile width G chosen for this model
  // override slot bookkeeping: (kind, id) -> (slot, ncomp)
  struct Slot { int kind, id, slot, ncomp; };
  std::vector<Slot> slots;
  std::vector<float> param0, init_qpos;
};

// returns "" on success; status gets a myo_status code on failure
std::string pack_model(const Model& m, const myo_task_cfg& cfg, PackedModel& out, int& status);
// every unsupported feature of the model, one "- ..." line each; empty when the model is inside the supported subset
std::string check_model(const Model& m);

}  // namespace myo

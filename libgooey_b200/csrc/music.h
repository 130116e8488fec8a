// music.h — host-side chord → MIDI-note tables behind gooey_engine_poly_trigger_chord (src/ffi.rs:5571-5611):
// Key::diatonic_sevenths (src/music/key.rs:55-84), ChordQuality::intervals (src/music/chord.rs:39-43),
// note_to_midi (src/music/note.rs:85-87) and apply_voicing (src/music/voicing.rs:76-172).  Integer logic only.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace gh {

// quality ids local to this file: 0 maj7, 1 m7, 2 dom7, 3 m7b5
inline const uint8_t* seventh_intervals(int q) {
  static const uint8_t T[4][4] = {{0, 4, 7, 11}, {0, 3, 7, 10}, {0, 4, 7, 10}, {0, 3, 6, 10}};
  return T[q];
}
inline uint8_t sat_sub12(uint8_t n) { return n >= 12 ? (uint8_t)(n - 12) : 0; }

// root 0-11 (id % 12), scale 0 major / 1 natural minor (anything else = major), degree % 7, voicing ids ffi.rs:5506-5515
// (unknown = root position), octave clamped to 0..8.  u8 additions wrap like the reference's release build would
// (debug builds panic on overflow; octave <= 8 keeps every sum below 256 except Spread, which saturates).
inline std::vector<uint8_t> chord_notes(uint32_t root, uint32_t scale, uint32_t degree, uint32_t voicing, int32_t octave) {
  static const uint8_t SCALE[2][7] = {{0, 2, 4, 5, 7, 9, 11}, {0, 2, 3, 5, 7, 8, 10}};
  static const uint8_t QUAL[2][7] = {{0, 1, 1, 0, 2, 1, 3}, {1, 3, 0, 1, 1, 0, 2}};
  const int sc = scale == 1 ? 1 : 0;
  const int deg = (int)(degree % 7u);
  const int root_idx = (int)((uint8_t)root % 12u);
  const int chord_root = (root_idx + SCALE[sc][deg]) % 12;
  const uint8_t* iv = seventh_intervals(QUAL[sc][deg]);
  const int oc = octave < 0 ? 0 : (octave > 8 ? 8 : octave);
  int rm = (oc + 1) * 12 + chord_root;
  const uint8_t root_midi = (uint8_t)(rm < 0 ? 0 : (rm > 127 ? 127 : rm));
  std::vector<uint8_t> n(4);
  for (int i = 0; i < 4; i++) n[i] = (uint8_t)(root_midi + iv[i]);
  switch (voicing) {
    case 1: n[0] += 12; std::sort(n.begin(), n.end()); break;
    case 2: n[0] += 12; n[1] += 12; std::sort(n.begin(), n.end()); break;
    case 3: n[0] += 12; n[1] += 12; n[2] += 12; std::sort(n.begin(), n.end()); break;
    case 4: for (size_t i = 1; i < n.size(); i += 2) n[i] += 12; std::sort(n.begin(), n.end()); break;
    case 5: n[2] = sat_sub12(n[2]); std::sort(n.begin(), n.end()); break;          // Drop2: second from the top
    case 6: break;                                                                // Drop3 needs >= 5 notes: unchanged
    case 7: for (size_t i = 0; i < n.size(); i++) { unsigned v = n[i] + (unsigned)(i / 2) * 12u; n[i] = (uint8_t)(v > 255u ? 255u : v); } std::sort(n.begin(), n.end()); break;
    case 8: n = {(uint8_t)(root_midi + iv[0]), (uint8_t)(root_midi + iv[1]), (uint8_t)(root_midi + iv[3])}; break;   // Shell: root, 3rd, 7th
    case 9: n.erase(n.begin()); n[0] = sat_sub12(n[0]); std::sort(n.begin(), n.end()); break;   // Rootless
    default: break;
  }
  for (auto& x : n) if (x > 127) x = 127;
  return n;
}

}  // namespace gh

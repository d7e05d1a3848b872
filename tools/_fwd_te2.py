import sys, torch
sys.path.insert(0,'/root/repo')
from oracle import decoder_oracle as O
from multi_speaker_tts_b200 import synthetic as S
from multi_speaker_tts_b200.decoder import decoder_forward
dev=torch.device('cuda:0')
for (B,Te,L) in [(7,160,5),(7,224,5),(16,256,6),(3,129,4),(20,200,4),(32,128,4)]:
    w=S.init_decoder_weights(0,bias_scale=0.05); b=S.synthetic_decoder_batch(B,Te,L,seed=B*100+Te,ragged=True)
    T=int(b['mel_len'].max())+1
    ref=O.decoder_forward(w,b['memory'],b['text_len'],b['mel'],b['mel_len'],b['prenet_mask'],b['zone_mask'])
    wd={k:v.to(dev) for k,v in w.items()}; bd={k:v.to(dev) for k,v in b.items()}
    lin,stop,align,_=decoder_forward(wd,bd['memory'],bd['text_len'],bd['mel'],bd['mel_len'],bd['prenet_mask'][:T].contiguous(),bd['zone_mask'][:T].contiguous(),is_training=True,n_steps=T,mode='bf16x3')
    torch.cuda.synchronize()
    e=[(lin.cpu()-ref[0]).abs().max().item(),(stop.cpu()-ref[1]).abs().max().item(),(align.cpu()-ref[2]).abs().max().item()]
    print(B,Te,L,'Linf',['%.2e'%x for x in e],'argmax eq',bool(torch.equal(align.cpu().argmax(-1),ref[2].argmax(-1))), 'text_len', b['text_len'].tolist()[:8])
